/*
 * manus_b200.h -- C ABI of the B200-native articulated-Gaussian-splat render path.
 *
 * Plain pointers and sizes only (no torch types).  Every pointer is a DEVICE pointer unless the
 * parameter name ends in _host.  All arrays are dense fp32 unless stated.  All entry points enqueue
 * work on `stream` and return immediately (0 = ok, non-zero = error; text via mb_last_error()).
 * The library is sm_100a only and has no CPU path.
 *
 * What each group replaces in the reference (paths relative to /root/reference):
 *
 *   mb_raster_*      the pybind module diff_gaussian_rasterization._C (third-party submodule, not in the tree;
 *                    installed by setup_env.sh:6,9-10) that GaussianRasterizer.forward/backward call:
 *                    rasterize_gaussians / rasterize_gaussians_backward / mark_visible.  Only call site in MANUS:
 *                    src/utils/gaussian_utils.py:378-416 (settings :378-391, call :407-416).
 *   mb_pose_*        the inline PyTorch pre-raster step, which has no operator boundary in the reference:
 *                    src/modules/hand_dynamic.py:86-137 (LBS of means and covariances),
 *                    src/models/gaussian.py:48-93 (activations, covariance build),
 *                    src/utils/gaussian_utils.py:248-314,431-449 + src/utils/sh_utils.py:57-120 (SH -> RGB).
 *   mb_dist2_knn3    simple_knn._C.distCUDA2 (third-party submodule, setup_env.sh:7,12-13); call site
 *                    src/models/gaussian.py:110.
 *
 * Matrix convention: viewmatrix / projmatrix are the 16 floats MANUS passes (world_view_transform and
 * full_proj_transform of src/utils/cam_utils.py:58-63, row-vector convention), i.e. element [4*col+row] of the
 * usual column-vector matrix.  bone_tf is [B,4,4] row-major column-vector transforms (x' = T x), as produced by
 * einsum("nij,njk->nik", posed, inv(rest)) in hand_dynamic.py:93-95.
 */
#ifndef MANUS_B200_H
#define MANUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *mb_stream_t; /* == cudaStream_t */

#define MB_OK 0
#define MB_ERR_INVALID 1   /* bad argument (shape / null / alignment) */
#define MB_ERR_CUDA 2      /* a CUDA runtime call or kernel launch failed */
#define MB_ERR_WORKSPACE 3 /* caller-provided workspace too small */

int mb_version(void);
const char *mb_last_error(void); /* thread-local, valid until the next call on this thread */
int mb_device_sm_count(void);    /* SM count of the current device (148 on B200); <0 on error */

/* Per-kernel timing with CUDA events on the launching stream (used by bench.py for the roofline numbers).
 * mb_profile_report synchronises the device and writes one line "kernel_name launches total_ms" per kernel launched
 * since the last report into buf (NUL terminated, truncated to cap). */
void mb_profile_enable(int on);
int mb_profile_report(char *buf, size_t cap);
/* One line "kernel_name stream start_us duration_us" per recorded launch, relative to the earliest start; the record is kept.
 * Launches recorded while a stream capture was active are event-record nodes of that graph and are re-recorded by every
 * replay: the timeline of a replayed multi-branch step (which kernels of which view overlap). */
int mb_profile_timeline(char *buf, size_t cap);

/* ------------------------------------------------------------------------------------------------
 * Rasterizer.  Mirrors RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible of the upstream
 * extension; the three opaque byte buffers play the role of upstream's
 * geomBuffer / binningBuffer / imgBuffer and must be kept alive (unmodified) from forward to backward.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mb_raster_inputs {
    int32_t num_points;       /* P */
    int32_t image_width;      /* GaussianRasterizationSettings.image_width  */
    int32_t image_height;     /* GaussianRasterizationSettings.image_height */
    int32_t sh_degree;        /* active degree D (0..3); used only when shs != NULL */
    int32_t sh_coeffs;        /* M = second dim of shs ([P,M,3]); 0 when shs == NULL */
    int32_t prefiltered;      /* accepted, unused (as upstream) */
    int32_t debug;            /* 1: synchronise + check after every stage */
    float tanfovx, tanfovy;
    float scale_modifier;
    const float *background;     /* [3]  */
    const float *means3D;        /* [P,3] */
    const float *opacities;      /* [P] (the [P,1] tensor) */
    const float *colors_precomp; /* [P,3] or NULL   } exactly one of the two */
    const float *shs;            /* [P,M,3] or NULL } */
    const float *cov3D_precomp;  /* [P,6] (xx,xy,xz,yy,yz,zz) or NULL } exactly one of */
    const float *scales;         /* [P,3] or NULL                      } cov3D_precomp | (scales & rotations) */
    const float *rotations;      /* [P,4] (r,x,y,z), used as given (not normalised) or NULL */
    const float *viewmatrix;     /* [16] */
    const float *projmatrix;     /* [16] */
    const float *campos;         /* [3]  */
    const float *tanfov_dev;     /* optional [2] = (tanfovx, tanfovy) in DEVICE memory; when non-NULL it replaces the two host
                                    values above, so that one enqueued frame (e.g. a captured CUDA graph) can be replayed with a
                                    different camera without touching kernel arguments */
} mb_raster_inputs;

size_t mb_raster_geom_bytes(int32_t num_points);
size_t mb_raster_binning_bytes(int64_t capacity, int32_t image_width, int32_t image_height);
size_t mb_raster_image_bytes(int32_t image_width, int32_t image_height);

/* Stage 1: per-Gaussian projection (cull, cov2D, conic, radius, tile count, SH->RGB when shs is given), depth
 * ordering and the instance offsets.  Writes radii[P] (0 for culled) and, if num_rendered_host != NULL (pinned host
 * memory), enqueues an async copy of num_rendered (= sum of tiles touched) into it.  No host synchronisation. */
int mb_raster_forward_geom(const mb_raster_inputs *in, void *geom, size_t geom_bytes, int32_t *radii,
                           int64_t *num_rendered_host, mb_stream_t stream);

/* Stage 2: instance emission, per-tile ordering, tile ranges, per-tile front-to-back alpha blend.
 * `capacity` = number of instances `binning` was sized for; if num_rendered > capacity nothing is written out of
 * bounds, the image is incomplete and the overflow is reported by mb_raster_query().  out_color is [3,H,W]. */
int mb_raster_forward_render(const mb_raster_inputs *in, void *geom, void *binning, size_t binning_bytes,
                             int64_t capacity, void *image_buf, size_t image_bytes, float *out_color, mb_stream_t stream);

/* Reads back (synchronises `stream`) the counters of the last forward held in `geom`:
 * num_rendered, number of visible Gaussians, overflow flag (1 if num_rendered exceeded capacity). */
int mb_raster_query(const void *geom, int64_t *num_rendered, int64_t *num_visible, int32_t *overflow, mb_stream_t stream);

/* Backward.  dL_dout is addressed as dL_dout[c*stride_c + y*stride_y + x*stride_x] (element strides), so both the
 * contiguous [3,H,W] tensor and the permuted view of an [H,W,3] tensor (src/utils/gaussian_utils.py:418) are read
 * in place.  Every output row is written (zeros for culled Gaussians); outputs for absent inputs may be NULL
 * (dL_dsh when shs == NULL; dL_dscales / dL_drotations when cov3D_precomp != NULL; dL_dcolors is written in both
 * colour modes).  grad_scratch: mb_raster_backward_scratch_bytes(P) bytes.  in->opacities is not read (may be NULL): the
 * opacities are part of the saved blend records, as in upstream, whose rasterize_gaussians_backward does not take them. */
size_t mb_raster_backward_scratch_bytes(int32_t num_points);
int mb_raster_backward(const mb_raster_inputs *in, const int32_t *radii, const void *geom, const void *binning,
                       int64_t capacity /* as given to forward_render */, const void *image_buf, const float *dL_dout, int64_t stride_c, int64_t stride_y, int64_t stride_x,
                       void *grad_scratch, size_t scratch_bytes, float *dL_dmeans2D /*[P,3]*/, float *dL_dcolors /*[P,3]*/,
                       float *dL_dopacity /*[P]*/, float *dL_dmeans3D /*[P,3]*/, float *dL_dcov3D /*[P,6]*/,
                       float *dL_dsh /*[P,M,3]*/, float *dL_dscales /*[P,3]*/, float *dL_drotations /*[P,4]*/,
                       mb_stream_t stream);

/* The tile half of the backward only: fills grad_scratch with one 12-float accumulator row per Gaussian: the pixel sums of
 * q = G dL/dalpha times (dx, dy, dx^2, dx dy, dy^2), d = mean2D - pixel (5) | sum q = dL/dopacity (1) | dL/dcolour (3) | 3 unused;
 * dL/dmean2D and dL/dconic follow from the five moments once per Gaussian (csrc/project_bwd.cuh).
 * mb_pose_backward_from_raster consumes the rows: on the fused path the
 * projection backward runs inside the pose backward kernel and dL_dmeans3D / dL_dcov3D / dL_dcolors / dL_dopacity never touch
 * HBM. */
int mb_raster_backward_blend(const mb_raster_inputs *in, const int32_t *radii, const void *geom, const void *binning,
                             int64_t capacity, const void *image_buf, const float *dL_dout, int64_t stride_c, int64_t stride_y,
                             int64_t stride_x, void *grad_scratch, size_t scratch_bytes, mb_stream_t stream);

/* Diagnostics for tests: byte offsets of named arrays inside the opaque buffers of a forward with these sizes.
 * out[0..3]: image buffer -> final_T f32[H*W], n_contrib u32[H*W], tile ranges u32[tiles][2], deepest last contributor u32[tiles];
 * out[4..5]: binning buffer -> gaussian id per sorted instance u32[num_rendered], tile id per sorted instance u32[num_rendered];
 * out[6..7]: geom buffer -> 48-B per-Gaussian blend records, counters u32[16]. */
int mb_raster_state_layout(int32_t num_points, int64_t capacity, int32_t image_width, int32_t image_height, int64_t *out,
                           int32_t n_out);

/* markVisible: out[i] = 1 iff the point is in front of the near plane (view-space z > 0.2). */
int mb_mark_visible(const float *means3D, int32_t num_points, const float *viewmatrix, const float *projmatrix,
                    uint8_t *out, mb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused pre-raster step (LBS + covariance + SH->RGB + activations), forward and backward.
 * Gaussians [0, num_skinned) are articulated (tf = sum_b w_b T_b); [num_skinned, num_points) are static (tf = I),
 * which covers hand-only (num_skinned = N), object-only (0) and composite scenes (src/modules/composite.py:50-60).
 * ---------------------------------------------------------------------------------------------- */
typedef struct mb_pose_inputs {
    int32_t num_points;
    int32_t num_skinned;
    int32_t num_bones;      /* B = columns of skin_wts = rows of bone_tf (identity "background" bone included) */
    int32_t sh_degree;      /* 0..3 */
    int32_t sh_coeffs;      /* K = 1 + second dim of f_rest; must be >= (sh_degree+1)^2 */
    int32_t isotropic;      /* 1: log_scale is [N,1] (opts.isotropic_scaling) */
    const float *xyz;           /* [N,3]  _xyz */
    const float *log_scale;     /* [N,3] or [N,1]  _scaling (pre-exp) */
    const float *quat;          /* [N,4]  _rotation (raw; normalised inside like build_rotation) */
    const float *opacity_logit; /* [N]    _opacity (pre-sigmoid) */
    const float *f_dc;          /* [N,1,3] _features_dc */
    const float *f_rest;        /* [N,K-1,3] _features_rest */
    const float *skin_wts;      /* [num_skinned,B] or NULL when num_skinned == 0 */
    const float *bone_tf;       /* [B,4,4] */
    const float *campos;        /* [3] */
    /* Optional: build the bone transforms inside the kernels (hand_dynamic.py:93-102) instead of passing bone_tf:
     * T_b = bones_posed[b] * bones_rest_inv[b] for b < num_posed_bones, identity for the other rows.  NULL = use bone_tf. */
    const float *bones_posed;    /* [num_posed_bones,4,4] */
    const float *bones_rest_inv; /* [num_posed_bones,4,4] inverse rest-pose transforms (constant per subject) */
    int32_t num_posed_bones;
    int32_t reserved_;
} mb_pose_inputs;

/* tf_out: optional [num_skinned,4,4] (the reference materialises it; the fused path does not need it). */
int mb_pose_forward(const mb_pose_inputs *in, float *posed_xyz /*[N,3]*/, float *posed_cov6 /*[N,6]*/,
                    float *colors /*[N,3]*/, float *opacity /*[N]*/, float *tf_out, mb_stream_t stream);

/* mb_pose_forward + the rasterizer's per-Gaussian forward (= mb_raster_forward_geom on the posed arrays) in one kernel: the
 * projection (A.1) runs on the posed mean / covariance / colour / opacity while they are in registers; geom / radii /
 * num_rendered_host exactly as mb_raster_forward_geom leaves them.  posed_xyz / posed_cov6 / colors / opacity are optional
 * (NULL: they never touch HBM; the fused backward mb_pose_backward_from_raster recomputes what it needs).  `raster` carries
 * the camera and image size (its per-Gaussian pointers are not read); continue with mb_raster_forward_render. */
int mb_pose_project_forward(const mb_pose_inputs *in, const mb_raster_inputs *raster, void *geom, size_t geom_bytes, int32_t *radii,
                            int64_t *num_rendered_host, float *posed_xyz, float *posed_cov6, float *colors, float *opacity,
                            mb_stream_t stream);

/* g_skin_wts: optional [num_skinned,B].  g_f_rest: optional (see mb_sh_grad_from_views).  All other outputs are required and
 * fully written. */
int mb_pose_backward(const mb_pose_inputs *in, const float *g_posed_xyz, const float *g_posed_cov6, const float *g_colors,
                     const float *g_opacity, float *g_xyz, float *g_log_scale, float *g_quat, float *g_opacity_logit,
                     float *g_f_dc, float *g_f_rest, float *g_skin_wts, mb_stream_t stream);

/* Pose backward fused with the rasterizer's projection backward (the part of RasterizeGaussiansBackwardCUDA after the tile
 * pass): `raster` is the mb_raster_inputs of the forward that rendered THIS pose's outputs
 * (colors_precomp / cov3D_precomp mode, scale_modifier 1), radii its radii, grad_scratch the rows written by
 * mb_raster_backward_blend.  Writes dL_dmeans2D [N,3] (the screen-space gradient MANUS's densification reads,
 * src/models/gaussian.py:335-338) and the parameter gradients; accumulate != 0 adds to them like mb_pose_backward_accumulate.
 * xyz_gradient_accum / denom / max_radii2D (all three or none): the densification statistics of add_densification_stats /
 * density_update (gaussian.py:335-338, gaussian_utils.py:461-473) updated in the same kernel for the visible Gaussians. */
int mb_pose_backward_from_raster(const mb_pose_inputs *in, const struct mb_raster_inputs *raster, const int32_t *radii,
                                 const void *grad_scratch, float *dL_dmeans2D, float *g_xyz, float *g_log_scale, float *g_quat,
                                 float *g_opacity_logit, float *g_f_dc, float *g_f_rest, float *g_skin_wts, int32_t accumulate,
                                 float *xyz_gradient_accum /*[N] or NULL*/, float *denom /*[N]*/, float *max_radii2D /*[N]*/,
                                 mb_stream_t stream);

/* The pose backward of ALL the views of one optimisation step in one pass (gradient accumulation over accum_iter views,
 * src/modules/hand_dynamic.py:248,259-277): the parameters of a tile are staged once, every thread walks the views of its Gaussian
 * (accumulator row, radius, pose and camera differ per view) and the summed parameter gradients are written once.  Same sums as
 * num_views calls of mb_pose_backward_from_raster with accumulate = 1, a quarter of the HBM traffic at four views.
 * `in`: the shared parameters (its bone_tf / bones_posed are ignored; bones_rest_inv / num_posed_bones are used with the views'
 * bones_posed).  Skin-weight gradients are not produced here. */
typedef struct mb_view_inputs {
    const struct mb_raster_inputs *raster; /* the view's forward: camera, image size */
    const int32_t *radii;                  /* [N] */
    const void *grad_scratch;              /* accumulator rows of the view's mb_raster_backward_blend */
    float *dL_dmeans2D;                    /* [N,3] out */
    const float *bone_tf;                  /* [B,4,4], or NULL with bones_posed */
    const float *bones_posed;              /* [num_posed_bones,4,4], or NULL with bone_tf */
    const float *campos;                   /* [3] */
} mb_view_inputs;
int mb_pose_backward_from_raster_views(const mb_pose_inputs *in, int32_t num_views, const mb_view_inputs *views, float *g_xyz,
                                       float *g_log_scale, float *g_quat, float *g_opacity_logit, float *g_f_dc, float *g_f_rest,
                                       int32_t accumulate, float *xyz_gradient_accum, float *denom, float *max_radii2D,
                                       mb_stream_t stream);

/* Same, but the gradients are ADDED to the output buffers (bulk TMA reduce-add, fp32 adds resolved in L2): gradient
 * accumulation over the views of one optimisation step -- the reference's accum_iter loop, src/modules/hand_dynamic.py:248,
 * 259-277, where autograd accumulates every view's gradient into the parameters' .grad.  The caller orders the
 * accumulating launches of one buffer (stream order or events) when it wants a deterministic summation order. */
int mb_pose_backward_accumulate(const mb_pose_inputs *in, const float *g_posed_xyz, const float *g_posed_cov6,
                                const float *g_colors, const float *g_opacity, float *g_xyz, float *g_log_scale, float *g_quat,
                                float *g_opacity_logit, float *g_f_dc, float *g_f_rest, float *g_skin_wts, mb_stream_t stream);

/* SH-coefficient gradients of a SUM over num_views views from the views' DC gradients (data-parallel step, SURVEY.md
 * section 8e): for one view g_f_rest[k][c] = basis_k(dir) * g_f_dc[c] / basis_0, with dir the canonical-space view direction
 * that mb_pose_forward uses (src/utils/gaussian_utils.py:431-449).  Ranks therefore all-gather g_f_dc (3 floats per Gaussian
 * and view) and the per-view (bone_tf, campos) instead of all-reducing the 48 SH floats, and call this on the gathered data.
 * bone_tf_all [num_views,B,4,4], campos_all [num_views,3], g_f_dc_all [num_views,N,3]; writes g_f_dc [N,3], g_f_rest [N,K-1,3].
 * view_stride = 0: the three per-view arrays are dense; otherwise view r of each starts view_stride floats after view r-1 (the
 * three pointers then address one gathered record per rank: DC gradients | bone transforms | camera centre). */
int mb_sh_grad_from_views(const float *xyz, int32_t num_points, const float *skin_wts, int32_t num_skinned, int32_t num_bones,
                          int32_t sh_degree, int32_t sh_coeffs, int32_t num_views, const float *bone_tf_all, const float *campos_all,
                          const float *g_f_dc_all, int64_t view_stride, float *g_f_dc, float *g_f_rest, mb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Per-frame skin-weight lookup: skin_wts[i,:] = normalise(trilinear(grid_weights, xyz[i])).
 * Replaces skinning_weights_from_voxel_grid (src/utils/gaussian_utils.py:167-196: grid_sample with align_corners=True and zero
 * padding on the [D,H,W,C] grid, coord = (xyz - grid_center) / grid_scale with x -> W, y -> H, z -> D, then w / w.sum(-1)),
 * called every step by HandGaussianModel.get_skin_weights (src/models/hand_gaussian.py:65-76).
 * grid_weights: [D,H,W,C] as stored by the reference (C <= 64); grid_center [3], grid_scale [3] in device memory.
 * Backward: g_xyz [N,3] is written; g_grid_weights ([D,H,W,C], may be NULL) is ACCUMULATED into (zero it first).
 * ---------------------------------------------------------------------------------------------- */
int mb_skin_weights_forward(const float *xyz, int32_t num_points, const float *grid_weights, int32_t depth, int32_t height,
                            int32_t width, int32_t channels, const float *grid_center, const float *grid_scale,
                            float *skin_wts /*[N,C]*/, mb_stream_t stream);
int mb_skin_weights_backward(const float *xyz, int32_t num_points, const float *grid_weights, int32_t depth, int32_t height,
                             int32_t width, int32_t channels, const float *grid_center, const float *grid_scale,
                             const float *g_skin_wts /*[N,C]*/, float *g_xyz /*[N,3]*/, float *g_grid_weights, mb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Photometric loss of the training step, value and gradient in one pass:
 *     loss = w_l1 * mean|pred - gt| + w_ssim * (1 - mean(ssim_map(pred, gt)))
 * Replaces l1_loss + ssim/_ssim of src/utils/loss_utils.py:22-97 as called by src/modules/base.py:323-365 on HWC
 * images (pred [H,W,3], gt [1,H,W,3]; weights config/{OBJ_GAUSSIAN,HAND_GAUSSIAN,COMPOSITE}.yaml:22-23).  Because the
 * reference takes `channel = img.size(-3)` (= H for HWC), its 11x11 SSIM window filters every image row separately over
 * the (W, 3) plane; this entry point reproduces exactly that.
 * gt, d_pred: dense [H,W,3].  pred is addressed as pred[y*stride_y + x*stride_x + c*stride_c] (element strides), so the
 * rasterizer's [3,H,W] output permuted to HWC (src/utils/gaussian_utils.py:418) is read in place.
 * loss_out[3] = (loss, mean |pred - gt|, mean ssim).  d_pred = d loss / d pred.
 * ---------------------------------------------------------------------------------------------- */
size_t mb_photometric_loss_workspace_bytes(int32_t height, int32_t width);
int mb_photometric_loss(const float *pred, int64_t pred_stride_y, int64_t pred_stride_x, int64_t pred_stride_c,
                        const float *gt, int32_t height, int32_t width, float w_l1, float w_ssim,
                        float *loss_out /*[3]*/, float *d_pred /*[H,W,3]*/, void *workspace, size_t workspace_bytes,
                        mb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused Adam over the flat parameter buffer.  Replaces the per-step work of
 * torch.optim.Adam(l, lr=0, eps=1e-15) with its six lr-only param groups (src/models/gaussian.py:133-141): one kernel over
 * elements [begin, end) of flat buffers whose segments [segment_end[s-1], segment_end[s]) use learning rate lr[s].
 * step = 1-based step count t (bias corrections 1 - beta^t as in torch); grad is multiplied by grad_scale first
 * (e.g. 1 / number of views).  segment_end_host / lr_host are HOST arrays of num_segments (<= 8) entries.
 * ---------------------------------------------------------------------------------------------- */
int mb_fused_adam(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t begin, int64_t end,
                  int32_t num_segments, const int64_t *segment_end_host, const double *lr_host, int64_t step, double beta1,
                  double beta2, double eps, float grad_scale, mb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * distCUDA2: mean squared distance to the 3 nearest other points (exact).
 * ---------------------------------------------------------------------------------------------- */
size_t mb_knn_workspace_bytes(int32_t num_points);
int mb_dist2_knn3(const float *points /*[N,3]*/, int32_t num_points, float *out /*[N]*/, void *workspace,
                  size_t workspace_bytes, mb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Contact distance: for every query point the Euclidean distance to, and the index of, its
 * nearest reference point (exact; ties -> lowest index).  Replaces get_contact_dist (src/utils/gaussian_utils.py:521-554, an
 * O(N*M) taichi loop; used by src/modules/composite.py:151-175 and scripts/process/mano_contacts.py:35) and get_contact_map
 * (:514-518).  workspace: mb_nearest_workspace_bytes(num_queries, num_refs).
 * ---------------------------------------------------------------------------------------------- */
size_t mb_nearest_workspace_bytes(int32_t num_queries, int32_t num_refs);
int mb_nearest_point(const float *queries /*[N,3]*/, int32_t num_queries, const float *refs /*[M,3]*/, int32_t num_refs,
                     float *out_dist /*[N]*/, int32_t *out_index /*[N]*/, void *workspace, size_t workspace_bytes, mb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Utility exported for tests: stable LSD radix sort of (u32 key, u32 value) pairs on key bits [0, end_bit).
 * n may be given on the host (n_host >= 0) or read from device memory (*n_dev, when n_host < 0, bounded by max_n).
 * Result lands in keys_out / vals_out.  workspace: mb_sort_workspace_bytes(max_n).
 * ---------------------------------------------------------------------------------------------- */
size_t mb_sort_workspace_bytes(int64_t max_n);
int mb_radix_sort_pairs(uint32_t *keys_in, uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out, int64_t n_host,
                        const uint32_t *n_dev, int64_t max_n, int32_t end_bit, void *workspace, size_t workspace_bytes,
                        mb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * calculate_colors_from_sh (src/utils/gaussian_utils.py:431-449) + eval_sh (src/utils/sh_utils.py:57-120) for callers that hold the
 * materialised per-Gaussian transform tf [N,4,4] (unchanged MANUS: render_gaussians(..., tf=...)): colour = max(SH(dir) + 0.5, 0)
 * with dir = normalize(means - (inv(tf) [campos; 1])[:3]) (tf given: means = canonical means) or normalize(means - campos)
 * (tf = NULL: means = posed means).  One thread per Gaussian, closed-form 4x4 inverse instead of torch.linalg.inv.
 * features: [N,sh_coeffs,3].  Backward writes g_means [N,3], g_features [N,sh_coeffs,3] and (optional) g_tf [N,4,4].
 * ---------------------------------------------------------------------------------------------- */
int mb_sh_colors_forward(const float *means, const float *features, const float *tf, const float *campos, int32_t num_points,
                         int32_t sh_degree, int32_t sh_coeffs, float *colors, mb_stream_t stream);
int mb_sh_colors_backward(const float *means, const float *features, const float *tf, const float *campos, int32_t num_points,
                          int32_t sh_degree, int32_t sh_coeffs, const float *g_colors, float *g_means, float *g_features, float *g_tf,
                          mb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * The gradient exchange of the view-sharded step over NVSwitch multicast memory (replaces the one ncclAllReduce per step,
 * SURVEY.md section 8e; reference semantics: the sum over accum_iter views, src/modules/hand_dynamic.py:248,259-277).
 * multicast_base: the multicast (multimem) address of a symmetric allocation that holds every rank's gradient buffer at the same
 * offsets.  pieces: up to 8 [offset, offset + count) float ranges of it; rank r reduces 1/world of every piece with
 * multimem.ld_reduce (fp32 add in the switch) and broadcasts the sums with multimem.st.  The caller synchronises the ranks before
 * (gradients complete everywhere) and after (stores landed everywhere).  max_ctas <= 0: one CTA per SM.
 * ---------------------------------------------------------------------------------------------- */
/* The same sum through plain peer-to-peer loads and stores on the R mapped buffers (peer_buffers[r] = rank r's copy): meant to run
 * BESIDE mb_multimem_allreduce on a disjoint part of every piece (the in-switch reduction does not saturate the links). */
int mb_p2p_allreduce(float *const *peer_buffers, const int64_t *piece_offsets, const int64_t *piece_counts, int32_t num_pieces,
                     int32_t rank, int32_t world, int32_t max_ctas, mb_stream_t stream);
int mb_multimem_allreduce(float *multicast_base, const int64_t *piece_offsets, const int64_t *piece_counts, int32_t num_pieces,
                          int32_t rank, int32_t world, int32_t max_ctas, mb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MANUS_B200_H */
