/*
 * isr.h -- C ABI of libisr.so, the B200 (sm_100a) surfel rasterizer for InstaScene.
 *
 * Drop-in boundary for the reference's native op  diff_surfel_rasterization._C  (DSR/ =
 * submodules/diff-surfel-rasterization) plus simple_knn._C.distCUDA2 and the sampled-pixel contrastive
 * loss.  Plain pointers and sizes only: no torch / glm / CUDA types in any signature.  All pointers are
 * DEVICE pointers unless the name ends in _host.  `stream` is a cudaStream_t passed as void*.
 * Every function returns 0 (ISR_OK) or a negative IsrStatus; nothing ever synchronises the device except
 * where stated.  The library owns no memory: outputs, gradients and workspaces are allocated by the
 * caller (DSR/rasterize_points.cu:88-109 does the same with torch tensors).
 *
 * Reference interface each entry point replaces (file:line relative to /root/reference):
 *   isr_forward_geometry + isr_forward_render  <- CudaRasterizer::Rasterizer::forward
 *                                                 DSR/cuda_rasterizer/rasterizer.h:31-58,
 *                                                 DSR/cuda_rasterizer/rasterizer_impl.cu:198-351,
 *                                                 bound as _C.rasterize_gaussians (DSR/ext.cpp:16,
 *                                                 DSR/rasterize_points.cu:39-151)
 *   isr_backward                               <- Rasterizer::backward  rasterizer.h:60-93,
 *                                                 rasterizer_impl.cu:355-463, bound as
 *                                                 _C.rasterize_gaussians_backward (DSR/ext.cpp:17,
 *                                                 DSR/rasterize_points.cu:153-262)
 *   isr_backward_extra_sparse                  <- same, restricted to dL/d(extra_attrs) for a list of
 *                                                 sampled pixels (the only non-zero cotangents in
 *                                                 train_semantic.py:118-141)
 *   isr_forward_sparse_extra / isr_backward_sparse_extra_views
 *                                              <- the same forward / backward restricted to the sampled pixels
 *                                                 of up to 8 views in one launch (train_semantic.py:102-173:
 *                                                 render() + mask gather + randint per term)
 *   isr_mark_visible                           <- Rasterizer::markVisible rasterizer.h:24-29,
 *                                                 rasterizer_impl.cu:141-153 (_C.mark_visible, ext.cpp:18)
 *   isr_geom_bytes / isr_image_bytes / isr_binning_bytes
 *                                              <- required<GeometryState|ImageState|BinningState>()
 *                                                 DSR/cuda_rasterizer/rasterizer_impl.h:29-72
 *   isr_contrastive_forward / _backward        <- utils/contrastive_utils.py:18-73 (contrastive_loss)
 *   isr_gather_pixels                          <- train_semantic.py:124-129 (boolean-mask gather + index)
 *   isr_sample_labelled                        <- train_semantic.py:118-126 (valid-pixel mask + randint)
 *   isr_aux_maps_forward / _backward           <- gaussian_renderer/__init__.py:127-156 + utils/point_utils.py:10-40
 *   isr_rownorm_forward / _backward            <- scene/gaussian_model.py:121-125 + gaussian_renderer/__init__.py:60-62
 *   isr_photometric_forward / _backward        <- l1_loss + ssim utils/loss_utils.py:18-19,39-83 as combined in
 *                                                 train.py:76-77
 *   isr_tracker_mark / isr_tracker_fill        <- get_segmap_gaussians spatial_track/modules/init_tracker.py:16-47
 *   isr_knn_mean_dist2                         <- SimpleKNN::knn submodules/simple-knn/simple_knn.cu:186-222
 *                                                 (distCUDA2, submodules/simple-knn/spatial.cu:15-25)
 */
#ifndef ISR_H_INCLUDED
#define ISR_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISR_VERSION 200

typedef enum IsrStatus {
    ISR_OK = 0,
    ISR_ERR_INVALID_ARG = -1,   /* null pointer / negative size / inconsistent optional inputs          */
    ISR_ERR_UNSUPPORTED = -2,   /* e.g. extra dims F > ISR_MAX_EXTRA_DIMS                                */
    ISR_ERR_WORKSPACE = -3,     /* a workspace is smaller than isr_*_bytes() says                        */
    ISR_ERR_CUDA = -4,          /* a CUDA call failed; isr_last_cuda_error() has the code                */
    ISR_ERR_NO_DEVICE = -5      /* no sm_100 device / kernels not loadable                                */
} IsrStatus;

#define ISR_MAX_EXTRA_DIMS 32   /* reference: MAX_EXTRA_DIMS 24 (DSR/cuda_rasterizer/auxiliary.h:20)      */
#define ISR_TILE 16             /* BLOCK_X = BLOCK_Y = 16 (DSR/cuda_rasterizer/config.h:16-17)            */

/* flags */
#define ISR_FLAG_BWD_WH_QUIRK 1u  /* backward uses W,H = int(focal*tan*2) like backward.cu:633-634 (Q5)   */
#define ISR_FLAG_NO_PAIRS     2u  /* do not emit the gau_related_pixels list                              */
#define ISR_FLAG_SKIP_BINNING 4u  /* isr_forward_render: reuse the binning already in the workspaces and run
                                     only the blend kernel (profiling / roofline measurement)              */
#define ISR_FLAG_SKIP_BLEND   16u  /* isr_forward_render: run only the binning (instance partition + tile ranges); the
                                     blend follows in a later call with ISR_FLAG_SKIP_BINNING on the same workspaces
                                     and the same R.  Binning reads no semantic features, so it can be queued ahead
                                     (instascene_b200.prefetch_geometry)                                       */
#define ISR_FLAG_SPEC_ARITH   8u  /* evaluate exp / rsqrt with the CPU-reproducible IEEE-only stand-ins of
                                     oracle/isr_oracle.c instead of CUDA's expf / rsqrtf (the reference's own,
                                     MUFU based).  Default (flag clear): the reference's arithmetic -- forward
                                     results are bit-identical to the unmodified reference CUDA rasterizer.
                                     Pass the same value to the forward and to the backward of a view.      */

/* gradient request mask for isr_backward (needs_input_grad gating; all = reference behaviour) */
#define ISR_GRAD_GEOMETRY 1u   /* means3D, means2D, scales, rotations, transMat, normal                  */
#define ISR_GRAD_COLOR    2u   /* colors_precomp / SH                                                    */
#define ISR_GRAD_OPACITY  4u
#define ISR_GRAD_EXTRA    8u
#define ISR_GRAD_ALL      15u

int isr_version(void);
const char* isr_status_string(int status);
int isr_last_cuda_error(void);              /* cudaError_t of the last failing CUDA call on this thread    */
int isr_device_sm_count(void);              /* SM count of the current device, or a negative IsrStatus    */
long long isr_kernel_launch_count(void);    /* kernels of THIS library launched by the process so far (CUB and
                                             * memset launches are not counted)                              */

/* ---- workspace sizes --------------------------------------------------------------------------------- */
size_t isr_geom_bytes(int P);               /* per-Gaussian state saved for backward                       */
size_t isr_image_bytes(int W, int H);       /* per-pixel state saved for backward + per-tile ranges        */
size_t isr_binning_bytes(int P, int64_t R, int W, int H);  /* instance list (per tile, depth order) + the partition's
                                                             * count table; R = capacity in instances, >= the
                                                             * emitted count num_rendered_host[1]              */

/* Offsets (in bytes) of the fields inside the geometry / image / binning workspaces, so that tests and
 * tools can compare intermediates with the oracle.  Field ids: */
enum IsrField {
    ISR_GEOM_SPLAT = 0,      /* float[P][16]: Tu[3] Tv[3] Tw[3] mean2D[2] normal[3] opacity, alpha-cut power  */
    ISR_GEOM_RGB = 1,        /* float[P][4]:  rgb, unused                                                 */
    ISR_GEOM_DEPTH = 2,      /* float[P]                                                                  */
    ISR_GEOM_TILES = 3,      /* uint32[P] tiles_touched                                                   */
    ISR_GEOM_CLAMPED = 4,    /* uint8[P]  bit c set <=> SH colour channel c was clamped                    */
    ISR_GEOM_DEPTH_ORDER = 5,/* uint32[P] Gaussian ids in ascending (depth bits, id) order                */
    ISR_GEOM_OFFSETS = 6,    /* uint32[P+1] exclusive scan of the emitted-tile counts in depth order      */
    ISR_GEOM_TILE_COUNT = 7, /* uint32[P] tiles actually emitted (<= tiles_touched): footprint-culled      */
    ISR_IMG_FINAL_T = 16,    /* float[3][H*W]: T, M1, M2                                                  */
    ISR_IMG_NCONTRIB = 17,   /* uint32[2][H*W]: last contributor, median contributor                      */
    ISR_IMG_RANGES = 18,     /* uint32[tiles][2]                                                          */
    ISR_BIN_POINT_LIST = 32  /* uint32[R] Gaussian ids sorted by (tile, depth bits, id): per tile the      *
                              * reference's list minus entries that provably reach none of its pixels     */
};
int64_t isr_field_offset(int field, int P, int64_t R, int W, int H);   /* <0: unknown field               */

/* ---- forward ----------------------------------------------------------------------------------------- */
typedef struct IsrForwardArgs {
    /* sizes */
    int P;                 /* number of Gaussians                                                         */
    int sh_degree;         /* active SH degree D (0..3)                                                   */
    int sh_coeffs;         /* M: coefficients per Gaussian in `shs` (0 if shs == NULL)                     */
    int F;                 /* extra (semantic feature) dims, 0..ISR_MAX_EXTRA_DIMS                        */
    int W, H;
    unsigned flags;
    /* camera */
    float tan_fovx, tan_fovy;
    float scale_modifier;
    const float* background;    /* [3]                                                                    */
    const float* viewmatrix;    /* [16] world_view_transform, m[4*col+row]                                */
    const float* projmatrix;    /* [16] full_proj_transform                                               */
    const float* campos;        /* [3]                                                                    */
    /* Gaussians (row-major fp32); exactly one of shs/colors_precomp and of (scales,rotations)/transMat_precomp */
    const float* means3D;       /* [P,3]                                                                  */
    const float* opacities;     /* [P]                                                                    */
    const float* scales;        /* [P,2] or NULL                                                          */
    const float* rotations;     /* [P,4] (w,x,y,z) or NULL                                                */
    const float* transMat_precomp; /* [P,9] or NULL                                                       */
    const float* shs;           /* [P,M,3] or NULL                                                        */
    const float* colors_precomp;/* [P,3] or NULL                                                          */
    const float* extra_attrs;   /* [P,F] or NULL when F == 0                                              */
    /* workspaces */
    void* geom;  size_t geom_bytes;
    void* image; size_t image_bytes;
    void* binning; size_t binning_bytes;      /* only needed by isr_forward_render                        */
    /* outputs */
    int* radii;                 /* [P] int32                                                              */
    float* out_color;           /* [3,H,W]                                                                */
    float* out_others;          /* [7,H,W]: depth, alpha, normal xyz, median depth, distortion            */
    float* out_extra;           /* [F,H,W] or NULL                                                        */
    int* pairs;                 /* [pair_capacity,2] (gaussian id, pixel id) or NULL                      */
    int64_t pair_capacity;      /* 9*H*W always suffices (sum of weights <= 1, each weight > 0.1)         */
    int* pair_count;            /* device int32: number of pairs written (NOT count-1 as in the reference)*/
    int64_t* num_rendered_host; /* pinned HOST int64[2] written asynchronously by isr_forward_geometry:   *
                                 * [0] = the reference's num_rendered (sum of tiles_touched),              *
                                 * [1] = R, the emitted instance count that sizes the binning workspace    */
} IsrForwardArgs;

/* Phase A: K1 preprocess (incl. per-Gaussian tile footprints) + depth ordering + offsets.  Enqueues an async
 * copy of the two instance counts into num_rendered_host[0..1]; the caller synchronises `stream`, sizes the
 * binning workspace with isr_binning_bytes(P, R = num_rendered_host[1], W, H) and calls phase B with that R.
 * (The reference blocks on a cudaMemcpy at the same point, rasterizer_impl.cu:287.) */
int isr_forward_geometry(const IsrForwardArgs* args, void* stream);
/* Phase B: stable partition of the instances by tile (fused with their emission) + tile ranges, then the front-to-back
 * blend.  R = the CAPACITY (in instances) of the binning workspace, >= num_rendered_host[1]; the kernels read the
 * actual count from device memory, so a caller may size the workspace from a high-water mark and queue phase B without
 * reading the count first (if it turns out too small nothing is written out of bounds; compare the count with R and
 * repeat phase B with a larger workspace). */
int isr_forward_render(const IsrForwardArgs* args, int64_t R, void* stream);

/* ---- backward ---------------------------------------------------------------------------------------- */
typedef struct IsrBackwardArgs {
    int P, sh_degree, sh_coeffs, F, W, H;
    unsigned flags;
    unsigned grad_mask;         /* ISR_GRAD_*                                                             */
    int64_t num_rendered;
    float tan_fovx, tan_fovy, scale_modifier;
    const float* background; const float* viewmatrix; const float* projmatrix; const float* campos;
    const float* means3D; const float* scales; const float* rotations; const float* transMat_precomp;
    const float* shs; const float* colors_precomp; const float* extra_attrs;
    const int* radii;
    const void* geom; const void* image; const void* binning;   /* as filled by the forward               */
    /* cotangents (CHW); NULL == all zeros */
    const float* dL_dcolor;     /* [3,H,W]                                                                */
    const float* dL_dothers;    /* [7,H,W]                                                                */
    const float* dL_dextra_pix; /* [F,H,W]                                                                */
    /* gradients; every requested buffer must be ZERO-FILLED by the caller (the reference binding does the
     * same with torch::zeros, rasterize_points.cu:207-220); NULL where not requested by grad_mask */
    float* dL_dmeans2D;   /* [P,3] */
    float* dL_dnormal;    /* [P,3] */
    float* dL_dopacity;   /* [P]   */
    float* dL_dcolors;    /* [P,3] */
    float* dL_dmeans3D;   /* [P,3] */
    float* dL_dtransMat;  /* [P,9] */
    float* dL_dsh;        /* [P,M,3] */
    float* dL_dscales;    /* [P,2] */
    float* dL_drotations; /* [P,4] */
    float* dL_dextra;     /* [P,F] */
} IsrBackwardArgs;

int isr_backward(const IsrBackwardArgs* args, void* stream);

/* dL/d(extra_attrs) only, for `n` sampled pixels: pix_ids[n] (= W*y+x, duplicates allowed) with cotangent
 * rows dL_dextra_samples[n,F].  Exactly equal (up to fp32 summation order) to isr_backward with a dense
 * [F,H,W] cotangent that is zero everywhere else.  dL_dextra [P,F] must be zero-filled by the caller. */
int isr_backward_extra_sparse(int P, int F, int W, int H, const float* extra_attrs, const void* geom,
                              const void* image, const void* binning, int64_t num_rendered, int n,
                              const int* pix_ids, const float* dL_dextra_samples, float* dL_dextra,
                              unsigned flags /* ISR_FLAG_SPEC_ARITH as in the forward */, void* stream);

/* ---- sampled-pixel rendering of the semantic features (train_semantic.py:102-173) ------------------------
 * The contrastive loop of the reference renders whole [F,H,W] feature maps (gaussian_renderer/__init__.py:101-113
 * through forward.cu:256-462) and then keeps `sample_batchsize` pixels of them (train_semantic.py:118-129; five
 * more full renders for the cross-view term, :146-173).  These two entry points composite ONLY the sampled
 * pixels, for up to ISR_MAX_SPARSE_VIEWS views of the same cloud in ONE launch: a warp per sample walks the
 * tile list of its pixel front to back with exactly the arithmetic and order of the dense blend (bit-identical
 * features), stops at saturation, and records n_contrib / final_T of that pixel for the backward.
 * A view must be PREPARED first: isr_forward_geometry + isr_forward_render(ISR_FLAG_SKIP_BLEND) into its own
 * geom / image / binning workspaces (the dense blend is never run).  All views share P, W, H. */
#define ISR_MAX_SPARSE_VIEWS 8
typedef struct IsrSparseView {
    const void* geom;     /* isr_geom_bytes(P), filled by isr_forward_geometry                              */
    void* image;          /* isr_image_bytes(W,H): tile ranges are read; n_contrib / final_T of the sampled
                             pixels are written by the forward and read by the backward                      */
    const void* binning;  /* isr_binning_bytes(...), filled by isr_forward_render(ISR_FLAG_SKIP_BLEND)      */
} IsrSparseView;

/* out_features[n,F] = rendered extra_attrs at pixel pix_ids[i] (= W*y+x, duplicates allowed) of view
 * view_ids[i] (NULL: every sample belongs to views_host[0]).  Samples with an id out of range give zeros. */
int isr_forward_sparse_extra(int n_views, const IsrSparseView* views_host, int P, int F, int W, int H,
                             const float* extra_attrs, int n, const int* pix_ids, const int* view_ids,
                             float* out_features, unsigned flags /* ISR_FLAG_SPEC_ARITH */, void* stream);

/* dL/d(extra_attrs) [P,F] (+=, zero-filled by the caller) from the cotangent rows dL_dfeatures[n,F] of the same
 * samples; needs the n_contrib entries the forward wrote (or those of a dense isr_forward_render). */
int isr_backward_sparse_extra_views(int n_views, const IsrSparseView* views_host, int P, int F, int W, int H, int n,
                                    const int* pix_ids, const int* view_ids, const float* dL_dfeatures,
                                    float* dL_dextra, unsigned flags, void* stream);

int isr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/* ---- sampled-pixel gather + ProtoNCE contrastive loss -------------------------------------------------- */
/* out[n,F] = feature_map[:, pix_ids[n]] for a CHW map [F, HW] */
int isr_gather_pixels(int F, int64_t HW, const float* feature_map, int n, const int* pix_ids, float* out,
                      void* stream);

/* n pixel ids drawn uniformly WITH replacement among the pixels whose label is > 0 (train_semantic.py:118-129 without
 * the boolean-mask gather of the feature map): sample i takes the labelled pixel of rank min(int(u[i] * n_valid),
 * n_valid - 1) in pixel order, u[i] in [0, 1) supplied by the caller (its random generator).  labels: [HW] signed
 * integers of label_bytes (1, 2, 4 or 8) bytes.  Writes pix_out[n] (int64) and labels_out[n] (int32, the label of each
 * drawn pixel).  No host synchronisation. */
size_t isr_sampler_workspace_bytes(int64_t HW);
int isr_sample_labelled(const void* labels, int label_bytes, int64_t HW, int n, const float* u, void* ws,
                        size_t ws_bytes, int64_t* pix_out, int* labels_out, void* stream);

size_t isr_contrastive_workspace_bytes(int N, int F, int K);
/* features [N,F], labels [N] int32 already shifted so that valid labels are 0..K-1 and invalid ones < 0;
 * predef_u [K,F] or NULL (cluster means).  min_pixnum: clusters with <= min_pixnum samples are dropped together with
 * their samples (utils/contrastive_utils.py:33-35; the reference's default is 0).  Writes *loss (device float) and saves
 * what backward needs in ws.  The similarity matrix runs on tcgen05 (3xTF32, fp32 accumulation in TMEM) in column
 * chunks of 256 clusters staged through shared memory: any K. */
int isr_contrastive_forward(int N, int F, int K, const float* features, const int* labels,
                            const float* predef_u, float temp_lambda, int min_pixnum, void* ws, size_t ws_bytes,
                            float* loss, void* stream);
/* dL_dfeatures[N,F] = grad_scale * d loss / d features */
int isr_contrastive_backward(int N, int F, int K, const float* features, const int* labels,
                             const float* predef_u, const void* ws, const float* grad_scale,
                             float* dL_dfeatures, void* stream);

/* ---- seg-feature activation ------------------------------------------------------------------------------ */
/* Row normalisation y = x / (|x| + eps1), optionally followed by a second one with eps2 (stages = 2): the two
 * L2 normalisations applied to _seg_feature before rasterisation (scene/gaussian_model.py:121-125 then
 * gaussian_renderer/__init__.py:60-62), fused into one pass each way.  x, y, dy, dx are [P,F]. */
int isr_rownorm_forward(int P, int F, const float* x, float eps1, float eps2, int stages, float* y, void* stream);
int isr_rownorm_backward(int P, int F, const float* x, const float* dy, float eps1, float eps2, int stages, float* dx,
                         void* stream);

/* ---- optimizer step of the trainable tensor ------------------------------------------------------------------- */
/* One fused pass of torch.optim.Adam's update (no weight decay / amsgrad / maximize) over n fp32 elements:
 * exp_avg = b1*exp_avg + (1-b1)*g; exp_avg_sq = b2*exp_avg_sq + (1-b2)*g*g;
 * param -= lr/(1-b1^step) * exp_avg / (sqrt(exp_avg_sq)/sqrt(1-b2^step) + eps).   `step` is the 1-based step count; with
 * step_dev != NULL the count is read from that DEVICE int32 instead (the caller increments it on the stream), so that a
 * captured CUDA graph replays with the right bias correction.
 * (scene/gaussian_model.py:217-249 trains _seg_feature with Adam(lr=0.025, eps=1e-15)) */
int isr_adam_step(size_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                  float beta2, float eps, int step, const int* step_dev, void* stream);
/* The same update for a [P,F] parameter whose gradient arrives as dy = dL/d(normalised rows) of isr_rownorm_forward
 * (same eps1, eps2, stages): the chain rule of isr_rownorm_backward is applied on the fly and the chained gradient is
 * never materialised (param, dy, exp_avg, exp_avg_sq are each read once).  grad_extra [P,F]: an ordinary gradient of the
 * parameter from its other uses, added after the chain rule, or NULL. */
int isr_adam_rownorm_step(int P, int F, float* param, const float* dy, const float* grad_extra, float* exp_avg,
                          float* exp_avg_sq, float eps1, float eps2, int stages, float lr, float beta1, float beta2, float eps,
                          int step, const int* step_dev, void* stream);

/* ---- derived maps of render() ------------------------------------------------------------------------------- */
/* Fused post-processing of allmap[7,H,W] (gaussian_renderer/__init__.py:127-156, utils/point_utils.py:10-40):
 * world-space normals, nan-cleaned median / expected depth, surf_depth and the finite-difference surf_normal.
 * normal_rot_host[9] (row-major M, n_world = n_view @ M = world_view_transform[:3,:3].T) and ray_mat_host[9]
 * (row-major K, ray = [x, y, 1] @ K) are HOST arrays (per-camera constants).  Outputs / gradient maps are CHW;
 * NULL gradient pointers mean zeros; g_allmap[7,H,W] is fully written. */
int isr_aux_maps_forward(int W, int H, const float* allmap, const float* normal_rot_host, const float* ray_mat_host,
                         float depth_ratio, float* rend_normal, float* rend_depth, float* rend_median, float* surf_depth,
                         float* surf_normal, void* stream);
int isr_aux_maps_backward(int W, int H, const float* allmap, const float* normal_rot_host, const float* ray_mat_host,
                          float depth_ratio, const float* g_rend_normal, const float* g_rend_depth,
                          const float* g_rend_median, const float* g_surf_depth, const float* g_surf_normal,
                          float* g_allmap, void* stream);

/* ---- simple-knn ------------------------------------------------------------------------------------- */
size_t isr_knn_workspace_bytes(int P);
int isr_knn_mean_dist2(int P, const float* points, float* mean_dist2, void* ws, size_t ws_bytes, void* stream);

/* ---- photometric loss of the RGB training step ------------------------------------------------ */
/* loss = (1 - lambda_dssim) * mean|image - gt| + lambda_dssim * (1 - SSIM(image, gt))  (train.py:76-77) with SSIM as
 * utils/loss_utils.py:39-83 (11x11 Gaussian window, sigma 1.5, zero padding, per channel, mean over C*H*W).
 * image, gt: [C,H,W] fp32.  forward writes out3 = {loss, L1 mean, SSIM mean} (device floats) and the derivative maps
 * backward needs into ws; backward writes dL_dimage[C,H,W] = *grad_scale (device float, NULL = 1) * d loss / d image. */
size_t isr_photometric_workspace_bytes(int C, int H, int W);
int isr_photometric_forward(int C, int H, int W, const float* image, const float* gt, float lambda_dssim, void* ws,
                            size_t ws_bytes, float* out3, void* stream);
int isr_photometric_backward(int C, int H, int W, const float* image, const float* gt, float lambda_dssim, const void* ws,
                             const float* grad_scale, float* dL_dimage, void* stream);

/* Densification statistics (train.py:139-142 + GaussianModel.add_densification_stats, scene/gaussian_model.py:602-605),
 * in place, for every Gaussian with radii > 0: max_radii2D = max(max_radii2D, radii); xyz_gradient_accum +=
 * |dL_dmeans2D[i, 0:3]|; denom += 1.   radii int32 [P], dL_dmeans2D fp32 [P,3], the three accumulators fp32 [P]. */
int isr_densify_stats(int P, const int* radii, const float* dL_dmeans2D, float* max_radii2D, float* xyz_gradient_accum,
                      float* denom, void* stream);

/* ---- Gaussian-tracker extraction --------------------------------------------------------- */
/* Device-side replacement of get_segmap_gaussians (spatial_track/modules/init_tracker.py:16-47): from the
 * (gaussian id, pixel id) pair list of one view and the view's segmentation map, the set of distinct Gaussians per
 * mask and of the whole frame -- without moving the pair list to the host.
 * seg_rows[HW] int32: per pixel the DENSE row of its mask, 0 = background / ignored, 1..K-1 = masks (the host mirror
 * maps the sorted distinct mask ids != 0 to rows 1..).  Row 0 of the result is the frame set.
 * isr_tracker_mark: builds one P-bit set per row in `ws` and writes counts[K] (device int32): distinct Gaussians per
 * row.  isr_tracker_fill: for every row with row_offsets[row] >= 0 (device int64[K]) writes the row's Gaussian ids in
 * ASCENDING order to out_ids[row_offsets[row] ...]; rows with a negative offset are skipped (the reference drops
 * masks with fewer than 50 Gaussians, init_tracker.py:41-42 -- that test runs on the host on K integers). */
size_t isr_tracker_workspace_bytes(int P, int K);
int isr_tracker_mark(const int* pairs, int64_t n_pairs, const int* seg_rows, int64_t HW, int P, int K, void* ws,
                     size_t ws_bytes, int* counts, void* stream);
int isr_tracker_fill(int P, int K, const void* ws, const int64_t* row_offsets, int* out_ids, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ISR_H_INCLUDED */
