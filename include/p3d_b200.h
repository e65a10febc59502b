/*
 * p3d_b200.h -- C ABI of libp3d_b200.so: the B200 (sm_100a) geometry hot path of
 * part-based 3-D reconstruction (orthographic semantic voxel carving + perspective
 * camera-candidate scoring).
 *
 * The reference (BarnitaSharma/Part-based-3D-Reconstruction) has no FFI; its
 * boundary is the set of Python functions the notebooks import.  Each entry point
 * below names the reference function (file:line under the reference root) whose
 * inner loop it replaces.  The Python layer in
 * part-based-3d-reconstruction_b200/utils/ keeps the reference signatures and calls
 * these through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the parameter is
 *     documented as "host"; the library never allocates persistent memory --
 *     scratch comes from caller-provided workspaces sized by *_workspace_bytes();
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*;
 *     NULL = the legacy default stream); the caller synchronises;
 *   - return value 0 = success, <0 = error (P3D_E_*); p3d_last_error() returns a
 *     thread-local message for the last failing call on this thread;
 *   - no global mutable state (host-side state such as helper streams lives in caller-owned
 *     context objects); re-entrant per stream;
 *   - arrays are C-order; grids are (A0,A1,A2) with flat index (a0*A1 + a1)*A2 + a2;
 *     RGB arrays carry a trailing channel axis of 3 uint8.
 */
#ifndef P3D_B200_H
#define P3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P3D_OK 0
#define P3D_E_INVALID (-1)   /* bad argument */
#define P3D_E_CUDA (-2)      /* CUDA runtime error (message has the cudaError string) */
#define P3D_E_WORKSPACE (-3) /* workspace too small */

typedef void* p3d_stream_t; /* cudaStream_t */

/* p3d_sweep mode */
#define P3D_MODE_JOINT 0    /* one label image per camera, last point in index order wins a pixel
                               (launch_smart_aligner: camera_estimation.py:552-572, 597-603) */
#define P3D_MODE_PER_PART 1 /* every part rendered on its own: pixel has part p iff any point of p lands
                               on it (visualize_voxel_projection_iou: camera_estimation.py:381-403) */

int p3d_version(void);
const char* p3d_last_error(void);
/* sm_count / cc_major / cc_minor / l2_bytes of the current device (host out-pointers, any may be NULL). */
int p3d_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes);

/* --------------------------------------------------------------------------------------------- *
 * Colour <-> label conversion (the reference compares RGB triples everywhere; the kernels work on
 * one-byte palette indices).
 * --------------------------------------------------------------------------------------------- */
/* labels[i] = 1 + (index of the first palette colour equal to rgb[i]), 0 if none.  n_colors <= 255.
 * Replaces the `np.all(x == colour, axis=-1)` scans at voxel_utils.py:12-15, mask_utils.py:92-95,
 * camera_estimation.py:777-779.  palette_rgb: device, n_colors*3 bytes. */
int p3d_rgb_to_labels(const uint8_t* rgb, int64_t n, const uint8_t* palette_rgb, int n_colors,
                      uint8_t* labels, p3d_stream_t stream);
/* rgb[i] = lut_rgb[labels[i]] ; lut_rgb: device, 256*3 bytes. */
int p3d_labels_to_rgb(const uint8_t* labels, int64_t n, const uint8_t* lut_rgb, uint8_t* rgb,
                      p3d_stream_t stream);

/* --------------------------------------------------------------------------------------------- *
 * get_voxel_points_by_parts            utils/voxel_utils.py:7-21
 * Stream compaction of the non-zero labels of a dense grid, in ascending flat index (the order of
 * np.where), as float32 points [x = a2, y = a1, z = a0].
 *   p3d_points_count : fills the workspace with per-tile offsets and writes the total to n_out[0]
 *   p3d_points_fill  : writes pts (n,3) f32 and pt_label (n) u8 using that workspace
 * --------------------------------------------------------------------------------------------- */
size_t p3d_points_workspace_bytes(int64_t n_voxels);
int p3d_points_count(const uint8_t* labels, int64_t n_voxels, int64_t* n_out, void* workspace,
                     size_t workspace_bytes, p3d_stream_t stream);
int p3d_points_fill(const uint8_t* labels, int A0, int A1, int A2, const void* workspace,
                    float* pts, uint8_t* pt_label, int64_t capacity, p3d_stream_t stream);

/* --------------------------------------------------------------------------------------------- *
 * look_at_rotation                     utils/camera_geometry.py:3-14
 * cand: (K,9) = cam_pos[3], target[3], f, cx, cy.  cams: (K,16) = cam_pos[3], R[9] row-major,
 * f, cx, cy, 0.  Same operation order as NumPy/OpenBLAS on the reference host.
 * --------------------------------------------------------------------------------------------- */
int p3d_setup_cameras_f64(const double* cand, int K, double* cams, p3d_stream_t stream);
int p3d_setup_cameras_f32(const float* cand, int K, float* cams, p3d_stream_t stream);

/* --------------------------------------------------------------------------------------------- *
 * project_colored_voxels (scatter)     utils/projection_utils.py:5-21
 * For each of K cameras and each point i: project, round half-even, bounds-test, then
 *   joint    : zbuf[k][v*W+u] = max(zbuf, i+1)        (last write wins == largest index wins)
 *   per-part : zbuf[k][v*W+u] |= 1 << (pt_label[i]-1) (pt_label required, values 1..32)
 * zbuf (K,H,W) uint32 must be zero on entry.  n < 2^32-1.
 * fast, bbox (both may be NULL): the (K,16) float blocks written by p3d_fast_cameras_f64 / _f32 for these cameras,
 * this image size and the bounding box `bbox` of `pts` (p3d_points_bbox; the kernel re-centres the points on the
 * middle of the box).  With them the kernel decides most pixels with a cheap FP32 evaluation under a proven error
 * bound and re-projects only the undecided points with the reference's exact sequence (float64 or float32, by entry
 * point) -- the z-buffer is bit-identical either way; without them every point takes the exact path.
 * --------------------------------------------------------------------------------------------- */
int p3d_splat_f64(const float* pts, const uint8_t* pt_label, int64_t n, const double* cams, int K,
                  int H, int W, int mode, uint32_t* zbuf, const float* fast, const float* bbox,
                  p3d_stream_t stream);
int p3d_splat_f32(const float* pts, const uint8_t* pt_label, int64_t n, const float* cams, int K,
                  int H, int W, int mode, uint32_t* zbuf, const float* fast, const float* bbox,
                  p3d_stream_t stream);

/* bbox (6 floats, device) = min x,y,z, max x,y,z of pts (n,3); NaN coordinates are ignored. */
int p3d_points_bbox(const float* pts, int64_t n, float* bbox, p3d_stream_t stream);
/* FP32 companion blocks of K cameras: pre-scaled rows, translation terms and the per-camera rounding thresholds
 * derived from an FP32 error bound over the box (see csrc/p3d_project.cuh); the _f32 variant adds the rounding error of
 * the float32 reference sequence itself to the bound.  fast: (K,16) floats. */
int p3d_fast_cameras_f64(const double* cams, int K, const float* bbox, int H, int W, float* fast,
                         p3d_stream_t stream);
int p3d_fast_cameras_f32(const float* cams, int K, const float* bbox, int H, int W, float* fast,
                         p3d_stream_t stream);

/* project_colored_voxels (image)       utils/projection_utils.py:20-23
 * img[p] = pt_rgb[zbuf[p]-1] or (0,0,0) where zbuf[p]==0.  One camera (joint-mode zbuf). */
int p3d_resolve_rgb(const uint32_t* zbuf, const uint8_t* pt_rgb, int64_t n_pixels, uint8_t* img,
                    p3d_stream_t stream);

/* --------------------------------------------------------------------------------------------- *
 * compute_partwise_iou                 utils/camera_estimation.py:770-787
 * --------------------------------------------------------------------------------------------- */
/* Two RGB images -> counts (P,2) int64 = (inter, union) per part colour.  part_rgb: device P*3. */
int p3d_partwise_counts_rgb(const uint8_t* proj_rgb, const uint8_t* gt_rgb, int64_t n_pixels,
                            const uint8_t* part_rgb, int P, int64_t* counts, p3d_stream_t stream);

/* --------------------------------------------------------------------------------------------- *
 * The candidate sweep: `evaluate` for K cameras in one call
 *                                      utils/camera_estimation.py:597-603 (+ :646 selection)
 *   pts (n,3) f32, pt_label (n) u8 in 1..P : the points of the selected parts, in index order
 *   cand (K,9)                            : candidate cameras (f64 or f32 arithmetic per entry point)
 *   gt_label (H,W) u8 in 0..P             : ground-truth part label per pixel (0 = none of the parts)
 *   counts (K,P,2) int64                  : (inter, union) per camera and part
 *   scores (K) f64                        : mean over the P parts of inter/union (0.0 if union == 0),
 *                                           summed in NumPy's pairwise order
 *   best (2) int64                        : [index of the first camera with the greatest score, 0]
 *                                           (may be NULL)
 * In P3D_MODE_PER_PART one extra row is appended per camera: counts is (K,P+1,2) and row P holds the
 * combined binary IoU counts against gt_any (H,W) u8 (camera_estimation.py:433-447); scores is the mean
 * over the P parts only.  gt_any may be NULL in joint mode.
 * Workspace: p3d_sweep_workspace_bytes(); cameras are processed in batches of z-buffers.
 *
 * segs / n_seg (optional, NULL / 0 = none): the x-run segments of `pts` from p3d_segments_fill().  With them the
 * sweep launches the segment splat (one thread per <= p3d_segment_length() voxels of a row at a constant x step, see
 * csrc/p3d_camera.cu); without them, for images above 2^22 pixels, or with P3D_SPLAT_POINTS=1, the per-point splat.
 * Both give identical counts.
 *
 * ctx (optional, NULL = none): caller-owned host state created by p3d_sweep_ctx_create() -- the helper stream and
 * events of the double-buffered batches, the launch counter and the optional splat timing.  With a context, sweeps of
 * more than one batch double-buffer the z-buffers and run the score pass of one batch on the context's helper stream
 * (forked from and joined to `stream` with events) beside the splat of the next; everything the caller enqueues on
 * `stream` afterwards is ordered after the whole sweep, also when the call fails midway.  Without a context, or while
 * `stream` is being captured into a CUDA graph, the batches stay in sequence on `stream`.  A context must not be used
 * by two calls at the same time; the library itself keeps no global mutable state.
 * --------------------------------------------------------------------------------------------- */
typedef struct p3d_sweep_ctx p3d_sweep_ctx;
p3d_sweep_ctx* p3d_sweep_ctx_create(void);              /* NULL on allocation failure */
void p3d_sweep_ctx_destroy(p3d_sweep_ctx* ctx);
/* kernel launches issued by the last p3d_sweep_* call that used ctx (bench accounting) */
int p3d_sweep_ctx_launches(const p3d_sweep_ctx* ctx);
/* Measurement: while enabled, p3d_sweep_* records a CUDA-event pair on the launch stream around every splat launch;
 * p3d_sweep_ctx_timing_read() waits for them, returns the summed splat duration (ms) and the number of launches
 * (host out-pointers), and resets the counters. */
int p3d_sweep_ctx_timing(p3d_sweep_ctx* ctx, int on);
int p3d_sweep_ctx_timing_read(p3d_sweep_ctx* ctx, double* splat_ms, int* n_launches);

size_t p3d_sweep_workspace_bytes(int K, int H, int W, int P, int elem_bytes);
int p3d_sweep_f64(const float* pts, const uint8_t* pt_label, int64_t n, const uint32_t* segs, int64_t n_seg,
                  const double* cand, int K, const uint8_t* gt_label, const uint8_t* gt_any, int H, int W, int P,
                  int mode, int64_t* counts, double* scores, int64_t* best, void* workspace,
                  size_t workspace_bytes, p3d_sweep_ctx* ctx, p3d_stream_t stream);
int p3d_sweep_f32(const float* pts, const uint8_t* pt_label, int64_t n, const uint32_t* segs, int64_t n_seg,
                  const float* cand, int K, const uint8_t* gt_label, const uint8_t* gt_any, int H, int W, int P,
                  int mode, int64_t* counts, double* scores, int64_t* best, void* workspace,
                  size_t workspace_bytes, p3d_sweep_ctx* ctx, p3d_stream_t stream);

/* x-run segments of a point list in get_voxel_points_by_parts order (utils/voxel_utils.py:17-19: ascending flat
 * index, so the voxels of one (z, y) row are consecutive in the list and in x).  A chunk = up to 32 L (L = seg_len)
 * list-consecutive points with equal label, equal (y, z), x increasing by exactly 1, not crossing a multiple of 32 L
 * in x.  A chunk of n points becomes T = ceil(n / L) segments; segment r owns the points r, r + T, r + 2T, ... of the
 * chunk (at most L).  Record = 4 uint32 { x_first | y << 16, z | (count-1) << 16 | (T-1) << 20 | label << 26,
 * list index of the first point, 0 }.
 *   p3d_segments_count : n_out (2) int64 device = [number of segments, number of points the segment form cannot
 *                        represent (non-integer or outside 0..65535, label outside 1..32)]; if n_out[1] != 0 the
 *                        caller must not pass segments to p3d_sweep_*.  workspace: p3d_segments_workspace_bytes(n).
 *   p3d_segments_fill  : writes the records (16-byte aligned, capacity >= n_out[0]) from the same workspace.
 * seg_len must equal p3d_segment_length() for segments handed to p3d_sweep_*. */
int p3d_segment_length(void);
size_t p3d_segments_workspace_bytes(int64_t n_points);
int p3d_segments_count(const float* pts, const uint8_t* pt_label, int64_t n, int seg_len, int64_t* n_out,
                       void* workspace, size_t workspace_bytes, p3d_stream_t stream);
int p3d_segments_fill(const float* pts, const uint8_t* pt_label, int64_t n, int seg_len, const void* workspace,
                      uint32_t* segs, int64_t capacity, p3d_stream_t stream);
/* Best-candidate reduction across GPUs (camera_estimation.py:646: the first candidate with the greatest
 * score wins).  p3d_best_pack writes pair = [bits of scores[best[0]], best[0] + offset] (index -1 if the
 * block was empty); the caller all-gathers the 16-byte pairs (NCCL); p3d_best_select picks the greatest
 * score with the lowest global index among n pairs into out (2) int64 = [score bits, index]. */
int p3d_best_pack(const double* scores, const int64_t* best, int64_t offset, int64_t* pair,
                  p3d_stream_t stream);
int p3d_best_select(const int64_t* pairs, int n, int64_t* out, p3d_stream_t stream);

/* --------------------------------------------------------------------------------------------- *
 * Depth-buffer visibility evaluator       utils/eval_helpers_intra.py:134-190  (SURVEY 8 f1, next to the path)
 *   p3d_depth_buffer_*  : compute_global_depth_buffer :134-161 -- zbuf (H,W) float32 = min Z of the points that project
 *                         into each pixel with Z > 1e-6, +inf elsewhere.  cam = one (16) camera block from
 *                         p3d_setup_cameras_*.  The f64 variant needs p3d_depth_workspace_bytes(H,W,8) of scratch
 *                         (64-bit atomicMin on the double's bits), the f32 variant none.
 *   p3d_part_visible_*  : project_part_visible :168-190 -- mask (H,W) uint8 = 1 where a point lies within eps of zbuf.
 * --------------------------------------------------------------------------------------------- */
size_t p3d_depth_workspace_bytes(int H, int W, int elem_bytes);
int p3d_depth_buffer_f32(const float* pts, int64_t n, const float* cam, int H, int W, float* zbuf, void* workspace,
                         size_t workspace_bytes, p3d_stream_t stream);
int p3d_depth_buffer_f64(const float* pts, int64_t n, const double* cam, int H, int W, float* zbuf, void* workspace,
                         size_t workspace_bytes, p3d_stream_t stream);
int p3d_part_visible_f32(const float* pts, int64_t n, const float* cam, const float* zbuf, float eps, int H, int W,
                         uint8_t* mask, p3d_stream_t stream);
int p3d_part_visible_f64(const float* pts, int64_t n, const double* cam, const float* zbuf, double eps, int H, int W,
                         uint8_t* mask, p3d_stream_t stream);

/* --------------------------------------------------------------------------------------------- *
 * Stage 3: part-wise deformation with a fixed camera      utils/deformation_estimation.py  
 * pts (n,3) float32 = voxel coordinates [x,y,z] of ONE part (get_voxel_points_by_parts); `stride` keeps every
 * stride-th point (project_fast :35-38).  deform = (scale_y, shift_y, scale_xz, shift_xz) doubles;
 * pix2vox = (W/W_img, H/H_img, D/W_img) doubles (:76-78).  All pointers are device pointers.
 *   p3d_deform_centres : sums (4) int64 = exact coordinate sums + number of non-integer coordinates (must be 0);
 *                        centres (9) doubles = per axis the mean for jitter offset 0, +0.25, -0.25 (:72, :87-96).
 *   p3d_deform_points  : deform_coords :70-98 before np.unique -- out (7, m, 3) int64, m = ceil(n/stride).
 *   p3d_pack_label_bits: bits[w] bit b = (labels[32w+b] == label), the ground-truth mask of one part.
 *   p3d_deform_sweep_* : for each of D candidates: deform, bounds-test in the (A0,A1,A2) grid (:111-115), project
 *                        through `cam` (one block of p3d_setup_cameras_f64/_f32 -- the working dtype of the
 *                        reference's projection follows the camera arrays; fast/bbox = its p3d_fast_cameras_f64/_f32
 *                        block for the box (0,0,0)-(A2-1,A1-1,A0-1), or both NULL = exact path only), and compare the
 *                        covered pixels with gt_bits:
 *                        counts (D,2) int64 = |proj & gt|, |proj | gt| (compute_partwise_iou for one part);
 *                        nvalid (D) int64 = (point, jitter) pairs inside the grid.  cov: (D, ceil(H*W/32)) uint32
 *                        scratch, zero on entry and on return.
 *   p3d_deform_scatter : save_deformed_grid :288-311 for one part -- grid_rgb (A0,A1,A2,3)[z][y][x] = (r,g,b) at
 *                        every valid deformed coordinate; nvalid (1) int64 may be NULL.
 * --------------------------------------------------------------------------------------------- */
int p3d_deform_centres(const float* pts, int64_t n, int64_t stride, int64_t* sums, double* centres,
                       p3d_stream_t stream);
int p3d_deform_points(const float* pts, int64_t n, int64_t stride, const double* centres, const double* deform,
                      const double* pix2vox, int64_t* out, p3d_stream_t stream);
int p3d_pack_label_bits(const uint8_t* labels, int64_t n, int label, uint32_t* bits, p3d_stream_t stream);
int p3d_deform_sweep_f64(const float* pts, int64_t n, int64_t stride, const double* centres, const double* deforms,
                         int D, const double* pix2vox, int A0, int A1, int A2, const double* cam, const float* fast,
                         const float* bbox, const uint32_t* gt_bits, int H, int W, uint32_t* cov, int64_t* counts,
                         int64_t* nvalid, p3d_stream_t stream);
int p3d_deform_sweep_f32(const float* pts, int64_t n, int64_t stride, const double* centres, const double* deforms,
                         int D, const double* pix2vox, int A0, int A1, int A2, const float* cam, const float* fast,
                         const float* bbox, const uint32_t* gt_bits, int H, int W, uint32_t* cov, int64_t* counts,
                         int64_t* nvalid, p3d_stream_t stream);
int p3d_deform_scatter(const float* pts, int64_t n, int64_t stride, const double* centres, const double* deform,
                       const double* pix2vox, int A0, int A1, int A2, int r, int g, int b, uint8_t* grid_rgb,
                       int64_t* nvalid, p3d_stream_t stream);

/* voxel_grid_to_points for RGB grids      utils/voxel_utils.py:35-51   (inspection without a CPU round trip)
 *   p3d_strided_occupancy  : mask (ceil(A0/s), ceil(A1/s), ceil(A2/s)) u8 = any(grid[::s, ::s, ::s], axis=-1)
 *   p3d_gather_scale_points: for the compacted points of that mask (p3d_points_fill: [x,y,z] on the sub-sampled
 *                            lattice) rgb[i] = grid[z*s][y*s][x*s] and pts[i] *= s (in place). */
int p3d_strided_occupancy(const uint8_t* grid_rgb, int A0, int A1, int A2, int stride, uint8_t* mask,
                          p3d_stream_t stream);
int p3d_gather_scale_points(const uint8_t* grid_rgb, int A0, int A1, int A2, int stride, float* pts, int64_t n,
                            uint8_t* rgb, p3d_stream_t stream);

/* meshify_colored_voxel_grid             utils/voxel_utils.py:53-95 (the marching-cubes call at :69-72)
 *   Marching cubes of a 0/1 volume `mask` (B0,B1,B2) at level 0.5 in this project's canonical order (scikit-image's
 *   Lewiner order is not reproduced: DESIGN.md 4.7): vertices = midpoints of the grid edges whose voxels differ, by (flat
 *   index of the lower voxel, axis); faces per cell in flat order from a 256-case table (ambiguous faces separate the
 *   occupied corners; counter-clockwise seen from the empty side); normals = normalised negated central-difference
 *   gradient.  p3d_mesh_count fills the workspace and writes totals (DEVICE, 2 x int64: vertices, faces); the caller
 *   reads them, allocates verts / normals (n_vertices,3) f32 in (a0,a1,a2) and faces (n_faces,3) i32, and calls
 *   p3d_mesh_emit with the same workspace. */
size_t p3d_mesh_workspace_bytes(int B0, int B1, int B2);
int p3d_mesh_count(const uint8_t* mask, int B0, int B1, int B2, void* workspace, size_t workspace_bytes, int64_t* totals,
                   p3d_stream_t stream);
int p3d_mesh_emit(const uint8_t* mask, int B0, int B1, int B2, void* workspace, size_t workspace_bytes, int64_t n_vertices,
                  int64_t n_faces, float* verts, float* normals, int32_t* faces, p3d_stream_t stream);

/* compute_binary_gt                      utils/eval_helpers_intra.py:274-285
 *   p3d_colour_presence: present (p3d_colour_presence_bytes() = 2 MiB, one bit per 24-bit colour r | g<<8 | b<<16)
 *                        = the non-black colours occurring in grid_rgb (n voxels).
 *   p3d_colour_lookup  : mask[p] = image colour p is marked in `present`. */
size_t p3d_colour_presence_bytes(void);
int p3d_colour_presence(const uint8_t* grid_rgb, int64_t n, uint32_t* present, p3d_stream_t stream);
int p3d_colour_lookup(const uint8_t* image_rgb, int64_t n, const uint32_t* present, uint8_t* mask,
                      p3d_stream_t stream);

/* ============================================================================================= *
 * Stage 1: orthographic semantic voxel carving            utils/voxel_carving_utils.py
 * Grids are (W,H,D) uint8 occupancy or (W,H,D,3) uint8 RGB.  2-D masks are passed already oriented:
 * mask_wh = (W,H) row-major (the result of the reference's _mask_to_wh :19-28), mask_hw = (H,W).
 * M (9 doubles, row-major) and off (3 doubles) are HOST pointers: the inverse rotation and offset that
 * the reference computes with NumPy (:65-69, :108, :119) and hands to scipy.ndimage.affine_transform.
 * ============================================================================================= */

/* One pass of process_voxel_grid :116-124 for any angle: scipy.ndimage.affine_transform(order=1,
 * mode="constant", cval=0) of a uint8 volume followed by the mask carve (mask_wh may be NULL). */
int p3d_resample_carve(const uint8_t* vol_in, int n0, int n1, int n2, const double* M, const double* off,
                       const uint8_t* mask_wh, uint8_t* vol_out, p3d_stream_t stream);

/* n_passes consecutive passes of the above (the cumulative rotations of process_voxel_grid :111-124 at angles that are
 * not index folds), ping-ponging between buf_a (holds the input) and buf_b; Ms (n_passes,9) / offs (n_passes,3) are HOST
 * arrays; *result_in_b (host) = 1 when the result ended up in buf_b. */
int p3d_resample_carve_passes(uint8_t* buf_a, uint8_t* buf_b, int n0, int n1, int n2, const double* Ms,
                              const double* offs, int n_passes, const uint8_t* mask_wh, int* result_in_b,
                              p3d_stream_t stream);

/* For a transform that leaves axis 1 alone: table (n0,n2) int32 = (src0 << 16 | src2) of the nearest source
 * voxel, -1 where scipy's bounds rule rejects the point; flag[0] != 0 when some in-range coordinate is not
 * within 1e-9 of an integer, i.e. the pass is NOT a pure index fold and needs p3d_resample_carve. */
int p3d_fold_table(int n0, int n2, const double* M, const double* off, int32_t* table, int* flag,
                   p3d_stream_t stream);
/* The same pass as p3d_resample_carve for a foldable transform: out = mask ? in[src0, y, src2] : 0. */
int p3d_fold_gather(const uint8_t* vol_in, int n0, int n1, int n2, const int32_t* table,
                    const uint8_t* mask_wh, uint8_t* vol_out, p3d_stream_t stream);

/* global_carve :269-298 at angle_interval = 90, fused with apply_colored_mask_to_voxel_grid :128-136:
 *   out[x,y,z] = colour[y,x] if mask[y,x] && table[x,z] >= 0 && mask[y, src0(x,z)] else 0
 * rgb != 0: colour_hw is (H,W,3) and out (W,H,D,3); rgb == 0: colour_hw is (H,W) labels and out (W,H,D). */
int p3d_global_carve_fold(int W, int H, int D, const int32_t* table, const uint8_t* mask_hw,
                          const uint8_t* colour_hw, int rgb, uint8_t* out, p3d_stream_t stream);
/* Bit-level form of p3d_global_carve_fold for z-separable tables (src0(x,z) = c - z on every in-range entry; true
 * for the 90-degree pass).  p3d_fold_analyse packs table >= 0 into inside_bits (W, ceil(D/32)) and writes
 * info (4 ints) = [max, min] of src0 + z and [max, min] of src2 - x over the in-range entries (z-separable iff
 * info[0] == info[1] =: c).  p3d_pack_mask_bits packs
 * mask_hw (H,W) into rows of words_per_row >= ceil(W/32) + 2 words with one zero word of padding on each side.
 * out must be 16-byte aligned; D a multiple of 32, or any D >= 16 with rgb != 0 (ragged rows: 16-voxel groups cut from
 * the flat voxel order, each spanning at most two z-rows).  Same output bytes as p3d_global_carve_fold.
 * [x_begin, x_begin + x_count) selects an x-slab: out is then the (x_count,H,D[,3]) slab -- the unit of multi-GPU
 * sharding (each output voxel depends only on the 2-D masks, so slabs need no exchange). */
int p3d_fold_analyse(const int32_t* table, int W, int D, uint32_t* inside_bits, int* info, p3d_stream_t stream);
int p3d_pack_mask_bits(const uint8_t* mask_hw, int H, int W, uint32_t* bits, int words_per_row,
                       p3d_stream_t stream);
int p3d_global_carve_fold_bits(int W, int H, int D, int x_begin, int x_count, const uint32_t* inside_bits, int c,
                               const uint32_t* mask_bits, int words_per_row, const uint8_t* colour_hw, int rgb,
                               uint8_t* out, p3d_stream_t stream);
/* carve_voxel_grid_with_masks :76-97: out = where(mask, grid, 0).  grid (W,H,D[,3]) with channels = 1 or 3;
 * mask_wh (W,H) with mask_channels = 1, or (W,H,3) with mask_channels = 3 (per-channel RGB branch :90-95). */
int p3d_mask_carve(const uint8_t* grid, int W, int H, int D, int channels, const uint8_t* mask_wh,
                   int mask_channels, uint8_t* out, p3d_stream_t stream);
/* apply_colored_mask_to_voxel_grid :128-136 (general path): out = colour[y,x,:] where carved == 1. */
int p3d_colourise(const uint8_t* carved, int W, int H, int D, const uint8_t* colour_hw, uint8_t* out,
                  p3d_stream_t stream);

/* part_carve :139-160 with every group at 90 degrees, all groups in one pass.  group_mask_hw (H,W) uint32:
 * bit g set when pixel (y,x) is in group g's mask AND in that mask after _mask_to_wh (square quirk). */
int p3d_part_carve_fold(const uint8_t* grid, int W, int H, int D, const int32_t* table,
                        const uint32_t* group_mask_hw, uint8_t* out, p3d_stream_t stream);

/* Bit-level form of p3d_part_carve_fold for z-separable tables with src0 = c - z and src2 = x + c2 (p3d_fold_analyse:
 * info[0] == info[1] =: c, info[2] == info[3] =: c2).  Needs D % 32 == 0 or D >= 16 (ragged rows), 16-byte aligned grids and
 * p3d_part_carve_bits_workspace_bytes() of scratch (z-packed occupancy and "alive" bits + per-group mask bits).
 * Two passes: the output is first written from the voxel-local terms, then the runs whose rotated source voxel is
 * empty are cleared (none for an already 4-way-symmetric grid). */
size_t p3d_part_carve_bits_workspace_bytes(int W, int H, int D, int n_groups);
int p3d_part_carve_fold_bits(const uint8_t* grid, int W, int H, int D, const uint32_t* inside_bits, int c, int c2,
                             const uint32_t* group_mask_hw, int n_groups, uint8_t* out, void* workspace,
                             size_t workspace_bytes, p3d_stream_t stream);
/* Output x slab [x_begin, x_begin + x_count) of the same carve (out_slab: (x_count,H,D,3)) from the whole, replicated
 * input grid -- the unit of multi-GPU sharding of part_carve (voxel_carving_utils.py:139-160).  No exchange between
 * ranks: besides its own rows the slab reads the z range [x_begin + c2, x_begin + x_count + c2) of every input row for
 * the rotated source occupancy.  Same workspace size as the full call. */
int p3d_part_carve_fold_bits_slab(const uint8_t* grid, int W, int H, int D, int x_begin, int x_count,
                                  const uint32_t* inside_bits, int c, int c2, const uint32_t* group_mask_hw,
                                  int n_groups, uint8_t* out_slab, void* workspace, size_t workspace_bytes,
                                  p3d_stream_t stream);
/* The same slab carve with a SHARDED input (every rank holds only its x slab): pass A writes out_slab and the slab's
 * rows of the z-packed occupancy bits -- the first W*H*(D/32) uint32 of the workspace, [x][y][word], slab rows
 * contiguous -- the ranks all-gather those rows (1/24 of the grid bytes; NCCL in utils/sweep.py), pass B clears the runs
 * whose rotated source is empty.  pass_a + pass_b over [0, W) equal p3d_part_carve_fold_bits. */
int p3d_part_carve_slab_pass_a(const uint8_t* grid_slab, int W, int H, int D, int x_begin, int x_count,
                               const uint32_t* inside_bits, int c, const uint32_t* group_mask_hw, int n_groups,
                               uint8_t* out_slab, void* workspace, size_t workspace_bytes, p3d_stream_t stream);
int p3d_part_carve_slab_pass_b(int W, int H, int D, int x_begin, int x_count, int c, int c2, uint8_t* out_slab,
                               void* workspace, size_t workspace_bytes, p3d_stream_t stream);
/* pass A split for callers that carve many grids with ONE mask and job list (the groups of voxel_carving_utils.py:143-146
 * do not depend on the grid): p3d_part_carve_pack_groups writes the per-group mask bits into the workspace once,
 * p3d_part_carve_slab_pass_a_packed is pass A without that step. */
int p3d_part_carve_pack_groups(const uint32_t* group_mask_hw, int W, int H, int D, int n_groups, void* workspace,
                               size_t workspace_bytes, p3d_stream_t stream);
int p3d_part_carve_slab_pass_a_packed(const uint8_t* grid_slab, int W, int H, int D, int x_begin, int x_count,
                                      const uint32_t* inside_bits, int c, const uint32_t* group_mask_hw, int n_groups,
                                      uint8_t* out_slab, void* workspace, size_t workspace_bytes, p3d_stream_t stream);
/* pass B fused with the exchange: instead of gathering the other ranks' occupancy rows first, the kernel reads every
 * source row straight from the workspace of the rank that owns it.  peer_workspaces: DEVICE array of n_ranks pointers,
 * entry r = rank r's workspace (the same layout on every rank) mapped into this process -- NVLink peer mappings, e.g.
 * torch symmetric memory; W must divide by n_ranks (rank r owns rows [r W/n, (r+1) W/n)).  The caller orders the call
 * after pass A of EVERY rank (any stream-ordered collective on `stream` does) and keeps the workspaces untouched until
 * every rank's pass B is done. */
int p3d_part_carve_slab_pass_b_peers(int W, int H, int D, int x_begin, int x_count, int c, int c2, uint8_t* out_slab,
                                     void* workspace, size_t workspace_bytes, const void* const* peer_workspaces,
                                     int n_ranks, p3d_stream_t stream);

/* Building blocks of the general-angle part_carve and of left_right_guided_carve :163-210. */
int p3d_crop_occupancy(const uint8_t* grid, int W, int H, int D, int x0, int y0, int z0, int w, int h, int d,
                       const uint8_t* sel_wh /* (w,h) or NULL */, uint8_t* occ, p3d_stream_t stream);
int p3d_accumulate_part(const uint8_t* grid, const uint8_t* carved, int W, int H, int D, const uint8_t* sel_wh,
                        uint8_t* final_grid, p3d_stream_t stream);
int p3d_paste_component(const uint8_t* src, const int32_t* labels, int comp, const uint8_t* kept, int W, int H,
                        int D, int x0, int y0, int z0, int w, int h, int d, uint8_t* out, p3d_stream_t stream);

/* left_right_guided_carve          utils/voxel_carving_utils.py:163-210, all components of one colour in one call:
 * crop every component's bounding box (occupancy of ANY colour, :191-193), run the n_pass rotate-and-carve passes of
 * process_voxel_grid (:104-126) on every crop at once (one launch per pass), paste (:199-201) and count the kept voxels
 * of each crop (the "carved voxels" log line, :196).
 *   labels (W,H,D) int32 from p3d_label6 of the colour's 3-D mask; mask_hw (H,W) u8 = the colour's 2-D mask (the crop
 *   mask is read in place: the reference's _mask_to_wh of an (h,w) crop is always its transpose);
 *   comps (n_comp,8) int32 device = { x0,y0,z0, w,h,d, scratch offset low, high } per component (1-based id = row + 1);
 *   Ms (n_pass,9), offs (n_comp,n_pass,3) doubles DEVICE: the host-computed inverse rotation of each pass and the
 *   offset for each crop shape (:116-123);  buf_a / buf_b: scratch of sum(w h d) bytes each;
 *   out: the result grid, a copy of `grid` on entry;  counts (n_comp) int64.
 *   sequential_paste != 0: one paste launch per component in id order (needed when bounding boxes overlap: the
 *   reference's loop order decides which component's write survives). */
int p3d_lr_carve_components(const uint8_t* grid, const int32_t* labels, int W, int H, int D, const uint8_t* mask_hw,
                            const int32_t* comps, int n_comp, int64_t max_crop_voxels, const double* Ms,
                            const double* offs, int n_pass, uint8_t* buf_a, uint8_t* buf_b, int sequential_paste,
                            uint8_t* out, int64_t* counts, p3d_stream_t stream);
/* part_carve's per-pixel group bits (:143-146 for every group at once): gm (H,W) uint32, bit g set where the pixel's
 * colour is one of group g's colours -- for a square image AND at the transposed pixel as well (the reference's
 * _mask_to_wh always transposes a square mask, so a group's effective mask is m & m.T).  keys[k] = r | g << 8 | b << 16,
 * group_of[k] = group of colour k (device arrays, n_keys <= 128). */
int p3d_group_image(const uint8_t* mask_rgb, int H, int W, const uint32_t* keys, const int32_t* group_of, int n_keys,
                    uint32_t* gm, p3d_stream_t stream);

/* np.all(grid == colour, axis=-1) :175,253 -> uint8 mask; a colour outside 0..255 matches nothing. */
int p3d_colour_mask(const uint8_t* grid_rgb, int64_t n, int r, int g, int b, uint8_t* mask, p3d_stream_t stream);

/* scipy.ndimage.label, default structure (6-connectivity) :175,254: labels (n0,n1,n2) int32 with ids
 * 1..n in raster order of each component's first voxel; n written to n_components[0] (device). */
size_t p3d_label6_workspace_bytes(int64_t n_voxels);
int p3d_label6(const uint8_t* mask, int n0, int n1, int n2, int32_t* labels, int32_t* n_components,
               void* workspace, size_t workspace_bytes, p3d_stream_t stream);
/* skimage.measure.label(mask) for a 2-D image (8-connectivity, ids in raster order of the first pixel):
 * camera_estimation.py:263 (extract_minaret_masks_by_label).  Workspace: p3d_label6_workspace_bytes(H*W). */
int p3d_label8_2d(const uint8_t* mask, int H, int W, int32_t* labels, int32_t* n_components, void* workspace,
                  size_t workspace_bytes, p3d_stream_t stream);
/* scipy.ndimage.label with structure = ones((3,3,3)) (26-connectivity; extract_top_k_components, voxel_utils.py:22-31);
 * same workspace and id order as p3d_label6. */
int p3d_label26(const uint8_t* mask, int n0, int n1, int n2, int32_t* labels, int32_t* n_components, void* workspace,
                size_t workspace_bytes, p3d_stream_t stream);
/* Per component: bbox (n,6) int32 = min0,min1,min2,max0,max1,max2 (inclusive) and sums (n,4) int64 =
 * voxel count and coordinate sums per axis (:184-185, :258-259). */
int p3d_component_stats(const int32_t* labels, int n0, int n1, int n2, int n_components, int32_t* bbox,
                        int64_t* sums, p3d_stream_t stream);
/* Top/bottom keypoints (camera_estimation.py:329-344): per component and end e (0 = min, 1 = max of the coordinate
 * along `axis`, read from bbox of p3d_component_stats), sums (n,2,4) int64 = count and coordinate sums of the voxels
 * AT that extreme. */
int p3d_component_extremes(const int32_t* labels, int n0, int n1, int n2, int n_components, int axis,
                           const int32_t* bbox, int64_t* sums, p3d_stream_t stream);
/* mask[i] = (labels[i] == id): one component as a 0/1 volume (np.argwhere(labeled == cid), camera_estimation.py:184). */
int p3d_label_equals(const int32_t* labels, int64_t n, int32_t id, uint8_t* mask, p3d_stream_t stream);
/* extract_top_bottom_voxel_points (camera_estimation.py:329-335) for a coordinate list (n,3) int32: minmax (2) int32 =
 * range of column `axis`; sums (2,4) int64 = count and column sums of the rows at the minimum / maximum. */
int p3d_coords_extremes(const int32_t* coords, int64_t n, int axis, int32_t* minmax, int64_t* sums,
                        p3d_stream_t stream);
/* recolor_backward_components :263-265: voxels of components with recolour[id-1] != 0 get the colour. */
int p3d_recolour_components(const int32_t* labels, const uint8_t* recolour, int64_t n, int r, int g, int b,
                            uint8_t* grid_rgb, p3d_stream_t stream);

/* extrude_from_surface :213-248, in place.  axis 2: mask_hw is (H,W) indexed [y][x]; axis 0: mask_hw is
 * (H,D) indexed [y][z].  sign = +1 / -1 for direction "+" / "-". */
int p3d_extrude(uint8_t* grid_rgb, int W, int H, int D, const uint8_t* mask_hw, int mask_h, int mask_w, int axis,
                int sign, int depth, int r, int g, int b, p3d_stream_t stream);

/* partwise_carve :384-385: out (D,H,W,3) = flip(transpose(in (W,H,D,3), (2,1,0,3)), axis=1). */
int p3d_reorient(const uint8_t* in, int W, int H, int D, uint8_t* out, p3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* P3D_B200_H */
