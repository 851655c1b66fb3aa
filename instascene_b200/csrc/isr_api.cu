// isr_api.cu -- the extern "C" boundary declared in include/isr.h.
#include <atomic>
#include <cstdio>

#include "isr_common.cuh"

namespace isr {
static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = e; }
static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int launch_preprocess_fwd(const IsrForwardArgs& a, cudaStream_t stream);
int launch_depth_order_and_offsets(const IsrForwardArgs& a, cudaStream_t stream);
int launch_binning(const IsrForwardArgs& a, int64_t R, cudaStream_t stream);
int launch_blend_fwd(const IsrForwardArgs& a, cudaStream_t stream);
int launch_blend_bwd(const IsrBackwardArgs& a, cudaStream_t stream);
int launch_preprocess_bwd(const IsrBackwardArgs& a, cudaStream_t stream);
int launch_extra_sparse_bwd(int P, int F, int W, int H, const void* geom, const void* image, const void* binning, int n,
                            const int* pix_ids, const float* dLdE, float* dL_dextra, unsigned flags, cudaStream_t stream);
int launch_sparse_bwd_views(int n_views, const IsrSparseView* views, int P, int F, int W, int H, int n, const int* pix_ids,
                            const int* view_ids, const float* dLdE, float* dL_dextra, unsigned flags, cudaStream_t stream);
int launch_sparse_fwd_views(int n_views, const IsrSparseView* views, int P, int F, int W, int H, const float* extras, int n,
                            const int* pix_ids, const int* view_ids, float* out, unsigned flags, cudaStream_t stream);
int launch_mark_visible(int P, const float* means3D, const float* view, const float* proj, uint8_t* present,
                        cudaStream_t stream);
size_t contrastive_ws_bytes(int N, int F, int K);
int launch_gather_pixels(int F, int64_t HW, const float* map, int n, const int* pix_ids, float* out, cudaStream_t stream);
int launch_contrastive_fwd(int N, int F, int K, const float* features, const int* labels, const float* predef_u,
                           float temp_lambda, int min_pixnum, void* ws, float* loss, cudaStream_t stream);
int launch_contrastive_bwd(int N, int F, int K, const int* labels, const float* predef_u, const void* ws,
                           const float* grad_scale, float* dfeat, cudaStream_t stream);
int launch_rownorm(bool fwd, int P, int F, const float* x, const float* dy, float e1, float e2, int stages, float* out,
                   cudaStream_t stream);
int launch_aux_fwd(int W, int H, const float* allmap, const float* M_host, const float* K_host, float depth_ratio,
                   float* rend_normal, float* rend_depth, float* rend_median, float* surf_depth, float* surf_normal,
                   cudaStream_t stream);
int launch_aux_bwd(int W, int H, const float* allmap, const float* M_host, const float* K_host, float depth_ratio,
                   const float* g_rend_normal, const float* g_rend_depth, const float* g_rend_median,
                   const float* g_surf_depth, const float* g_surf_normal, float* g_allmap, cudaStream_t stream);
int launch_adam(size_t n, float* p, const float* g, float* m, float* v, float lr, float beta1, float beta2, float eps,
                int step, const int* step_dev, cudaStream_t stream);
int launch_adam_rownorm(int P, int F, float* x, const float* dy, const float* g_extra, float* m, float* v, float e1, float e2,
                        int stages, float lr, float beta1, float beta2, float eps, int step, const int* step_dev,
                        cudaStream_t stream);
size_t sampler_ws_bytes(int64_t HW);
int launch_sampler(const void* labels, int label_bytes, int64_t HW, int n, const float* u, void* ws, int64_t* pix_out,
                   int* lab_out, cudaStream_t stream);
size_t knn_ws_bytes(int P);
int launch_knn(int P, const float* points, float* out, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t photometric_ws_bytes(int C, int H, int W);
int launch_photometric_fwd(int C, int H, int W, const float* img, const float* gt, float lambda, void* ws, float* out,
                           cudaStream_t stream);
int launch_photometric_bwd(int C, int H, int W, const float* img, const float* gt, float lambda, const void* ws,
                           const float* grad_scale, float* dimg, cudaStream_t stream);
int launch_densify_stats(int P, const int* radii, const float* grad2d, float* max_radii2D, float* grad_accum, float* denom,
                         cudaStream_t stream);
size_t tracker_ws_bytes(int P, int K);
int launch_tracker_mark(const int* pairs, int64_t n_pairs, const int* seg_rows, int64_t HW, int P, int K, void* ws,
                        int* counts, cudaStream_t stream);
int launch_tracker_fill(int P, int K, const void* ws, const int64_t* row_offsets, int* out_ids, cudaStream_t stream);
}  // namespace isr

using namespace isr;

extern "C" {

int isr_version(void) { return ISR_VERSION; }

const char* isr_status_string(int status) {
    switch (status) {
        case ISR_OK: return "ok";
        case ISR_ERR_INVALID_ARG: return "invalid argument";
        case ISR_ERR_UNSUPPORTED: return "unsupported configuration";
        case ISR_ERR_WORKSPACE: return "workspace too small";
        case ISR_ERR_CUDA: return "CUDA error";
        case ISR_ERR_NO_DEVICE: return "no usable CUDA device";
        default: return "unknown status";
    }
}

int isr_last_cuda_error(void) { return (int)g_last_cuda_error; }

long long isr_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int isr_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return ISR_ERR_NO_DEVICE;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return ISR_ERR_NO_DEVICE;
    return n;
}

size_t isr_geom_bytes(int P) { return GeomLayout(P).total + 256; }
size_t isr_image_bytes(int W, int H) { return ImageLayout(W, H).total + 256; }
size_t isr_binning_bytes(int P, int64_t R, int W, int H) { return BinLayout(P, R, W, H).total + 256; }

int64_t isr_field_offset(int field, int P, int64_t R, int W, int H) {
    switch (field) {
        case ISR_GEOM_SPLAT: return (int64_t)GeomLayout(P).splat;
        case ISR_GEOM_RGB: return (int64_t)GeomLayout(P).rgb;
        case ISR_GEOM_DEPTH: return (int64_t)GeomLayout(P).depth;
        case ISR_GEOM_TILES: return (int64_t)GeomLayout(P).tiles;
        case ISR_GEOM_CLAMPED: return (int64_t)GeomLayout(P).clamped;
        case ISR_GEOM_DEPTH_ORDER: return (int64_t)GeomLayout(P).order;
        case ISR_GEOM_OFFSETS: return (int64_t)GeomLayout(P).offsets;
        case ISR_GEOM_TILE_COUNT: return (int64_t)GeomLayout(P).tcount;
        case ISR_IMG_FINAL_T: return (int64_t)ImageLayout(W, H).final_T;
        case ISR_IMG_NCONTRIB: return (int64_t)ImageLayout(W, H).n_contrib;
        case ISR_IMG_RANGES: return (int64_t)ImageLayout(W, H).ranges;
        case ISR_BIN_POINT_LIST: return (int64_t)BinLayout(P, R, W, H).point_list;
        default: return -1;
    }
}

static int check_forward_args(const IsrForwardArgs* a) {
    if (!a) return ISR_ERR_INVALID_ARG;
    if (a->P < 0 || a->W <= 0 || a->H <= 0 || a->F < 0) return ISR_ERR_INVALID_ARG;
    if (a->F > ISR_MAX_EXTRA_DIMS) return ISR_ERR_UNSUPPORTED;
    if (a->P == 0) return ISR_OK;
    if (!a->means3D || !a->opacities || !a->viewmatrix || !a->projmatrix || !a->background) return ISR_ERR_INVALID_ARG;
    // exactly one of SH / precomputed colours, exactly one of (scales, rotations) / transMat_precomp
    if ((a->shs == nullptr) == (a->colors_precomp == nullptr)) return ISR_ERR_INVALID_ARG;
    const bool has_sr = a->scales != nullptr && a->rotations != nullptr;
    if (has_sr == (a->transMat_precomp != nullptr)) return ISR_ERR_INVALID_ARG;
    if (!has_sr && (a->scales != nullptr || a->rotations != nullptr)) return ISR_ERR_INVALID_ARG;
    if (a->shs && (!a->campos || a->sh_coeffs < (a->sh_degree + 1) * (a->sh_degree + 1) || a->sh_degree < 0 || a->sh_degree > 3))
        return ISR_ERR_INVALID_ARG;
    if (a->F > 0 && (!a->extra_attrs || !a->out_extra)) return ISR_ERR_INVALID_ARG;
    if (!a->geom || !a->image || !a->radii || !a->out_color || !a->out_others) return ISR_ERR_INVALID_ARG;
    if (a->geom_bytes < GeomLayout(a->P).total || a->image_bytes < ImageLayout(a->W, a->H).total) return ISR_ERR_WORKSPACE;
    return ISR_OK;
}

int isr_forward_geometry(const IsrForwardArgs* a, void* stream_) {
    int st = check_forward_args(a);
    if (st != ISR_OK) return st;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (a->P == 0) {
        if (a->num_rendered_host) a->num_rendered_host[0] = a->num_rendered_host[1] = 0;
        return ISR_OK;
    }
    st = launch_preprocess_fwd(*a, stream);
    if (st != ISR_OK) return st;
    return launch_depth_order_and_offsets(*a, stream);
}

int isr_forward_render(const IsrForwardArgs* a, int64_t R, void* stream_) {
    int st = check_forward_args(a);
    if (st != ISR_OK) return st;
    if (R < 0 || R > 0x7fffffffLL) return ISR_ERR_INVALID_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool do_bin = !(a->flags & ISR_FLAG_SKIP_BINNING), do_blend = !(a->flags & ISR_FLAG_SKIP_BLEND);
    if (a->P > 0 && R > 0) {
        if (!a->binning || a->binning_bytes < BinLayout(a->P, R, a->W, a->H).total) return ISR_ERR_WORKSPACE;
    }
    if (do_bin) {
        st = launch_binning(*a, a->P > 0 ? R : 0, stream);
        if (st != ISR_OK) return st;
    }
    if (!do_blend) return ISR_OK;
    if (a->F > 0 && (!a->extra_attrs || !a->out_extra)) return ISR_ERR_INVALID_ARG;
    if (a->pair_count) ISR_CUDA_TRY(cudaMemsetAsync(a->pair_count, 0, sizeof(int), stream));
    // With no instances every tile range is (0,0): the blend kernel still runs to write background / zeros.
    return launch_blend_fwd(*a, stream);
}

int isr_backward(const IsrBackwardArgs* a, void* stream_) {
    if (!a) return ISR_ERR_INVALID_ARG;
    if (a->P < 0 || a->W <= 0 || a->H <= 0 || a->F < 0) return ISR_ERR_INVALID_ARG;
    if (a->F > ISR_MAX_EXTRA_DIMS) return ISR_ERR_UNSUPPORTED;
    if (a->P == 0) return ISR_OK;
    if (!a->geom || !a->image || !a->radii || !a->means3D || !a->viewmatrix || !a->projmatrix || !a->background)
        return ISR_ERR_INVALID_ARG;
    if (a->num_rendered > 0 && !a->binning) return ISR_ERR_INVALID_ARG;
    const unsigned m = a->grad_mask;
    if ((m & ISR_GRAD_GEOMETRY) && (!a->dL_dmeans2D || !a->dL_dnormal || !a->dL_dtransMat || !a->dL_dmeans3D))
        return ISR_ERR_INVALID_ARG;
    if ((m & ISR_GRAD_GEOMETRY) && a->scales && (!a->dL_dscales || !a->dL_drotations)) return ISR_ERR_INVALID_ARG;
    if ((m & ISR_GRAD_COLOR) && !a->dL_dcolors) return ISR_ERR_INVALID_ARG;
    if ((m & ISR_GRAD_OPACITY) && !a->dL_dopacity) return ISR_ERR_INVALID_ARG;
    if ((m & ISR_GRAD_EXTRA) && a->F > 0 && !a->dL_dextra) return ISR_ERR_INVALID_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int st = launch_blend_bwd(*a, stream);
    if (st != ISR_OK) return st;
    // K8 needs the geometry-gradient buffers (it also produces the SH gradient from dL_dcolors)
    if ((m & ISR_GRAD_GEOMETRY) && (m & ISR_GRAD_COLOR)) return launch_preprocess_bwd(*a, stream);
    if (m & (ISR_GRAD_GEOMETRY | ISR_GRAD_COLOR)) {
        // partial requests: run K8 only when everything it touches is present
        if (a->dL_dmeans2D && a->dL_dnormal && a->dL_dtransMat && a->dL_dmeans3D && a->dL_dcolors)
            return launch_preprocess_bwd(*a, stream);
    }
    return ISR_OK;
}

int isr_backward_extra_sparse(int P, int F, int W, int H, const float* extra_attrs, const void* geom, const void* image,
                              const void* binning, int64_t num_rendered, int n, const int* pix_ids,
                              const float* dL_dextra_samples, float* dL_dextra, unsigned flags, void* stream_) {
    (void)extra_attrs;
    if (P < 0 || F < 0 || W <= 0 || H <= 0 || n < 0) return ISR_ERR_INVALID_ARG;
    if (F > ISR_MAX_EXTRA_DIMS) return ISR_ERR_UNSUPPORTED;
    if (P == 0 || n == 0 || F == 0 || num_rendered <= 0) return ISR_OK;
    if (!geom || !image || !binning || !pix_ids || !dL_dextra_samples || !dL_dextra) return ISR_ERR_INVALID_ARG;
    return launch_extra_sparse_bwd(P, F, W, H, geom, image, binning, n, pix_ids, dL_dextra_samples, dL_dextra, flags,
                                   static_cast<cudaStream_t>(stream_));
}

static int check_sparse_views(int n_views, const IsrSparseView* views, int P, int F, int W, int H, int n) {
    if (P < 0 || F < 0 || W <= 0 || H <= 0 || n < 0 || n_views < 0) return ISR_ERR_INVALID_ARG;
    if (F > ISR_MAX_EXTRA_DIMS || n_views > ISR_MAX_SPARSE_VIEWS) return ISR_ERR_UNSUPPORTED;
    if (n > 0 && P > 0 && F > 0) {
        if (n_views < 1 || !views) return ISR_ERR_INVALID_ARG;
        for (int i = 0; i < n_views; i++)
            if (!views[i].geom || !views[i].image || !views[i].binning) return ISR_ERR_INVALID_ARG;
    }
    return ISR_OK;
}

int isr_forward_sparse_extra(int n_views, const IsrSparseView* views_host, int P, int F, int W, int H, const float* extra_attrs,
                             int n, const int* pix_ids, const int* view_ids, float* out_features, unsigned flags,
                             void* stream_) {
    const int st = check_sparse_views(n_views, views_host, P, F, W, H, n);
    if (st != ISR_OK) return st;
    if (n == 0 || F == 0) return ISR_OK;
    if (!pix_ids || !out_features) return ISR_ERR_INVALID_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (P == 0) {  // nothing to composite: every sample renders zeros
        ISR_CUDA_TRY(cudaMemsetAsync(out_features, 0, sizeof(float) * (size_t)n * F, stream));
        return ISR_OK;
    }
    if (!extra_attrs) return ISR_ERR_INVALID_ARG;
    return launch_sparse_fwd_views(n_views, views_host, P, F, W, H, extra_attrs, n, pix_ids, view_ids, out_features, flags, stream);
}

int isr_backward_sparse_extra_views(int n_views, const IsrSparseView* views_host, int P, int F, int W, int H, int n,
                                    const int* pix_ids, const int* view_ids, const float* dL_dfeatures, float* dL_dextra,
                                    unsigned flags, void* stream_) {
    const int st = check_sparse_views(n_views, views_host, P, F, W, H, n);
    if (st != ISR_OK) return st;
    if (n == 0 || F == 0 || P == 0) return ISR_OK;
    if (!pix_ids || !dL_dfeatures || !dL_dextra) return ISR_ERR_INVALID_ARG;
    return launch_sparse_bwd_views(n_views, views_host, P, F, W, H, n, pix_ids, view_ids, dL_dfeatures, dL_dextra, flags,
                                   static_cast<cudaStream_t>(stream_));
}

int isr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                     void* stream_) {
    if (P < 0) return ISR_ERR_INVALID_ARG;
    if (P == 0) return ISR_OK;
    if (!means3D || !viewmatrix || !present) return ISR_ERR_INVALID_ARG;
    return launch_mark_visible(P, means3D, viewmatrix, projmatrix, present, static_cast<cudaStream_t>(stream_));
}

int isr_gather_pixels(int F, int64_t HW, const float* feature_map, int n, const int* pix_ids, float* out, void* stream_) {
    if (F < 0 || n < 0 || HW < 0) return ISR_ERR_INVALID_ARG;
    if (F == 0 || n == 0) return ISR_OK;
    if (!feature_map || !pix_ids || !out) return ISR_ERR_INVALID_ARG;
    return launch_gather_pixels(F, HW, feature_map, n, pix_ids, out, static_cast<cudaStream_t>(stream_));
}

size_t isr_sampler_workspace_bytes(int64_t HW) { return HW < 0 ? 0 : sampler_ws_bytes(HW); }

int isr_sample_labelled(const void* labels, int label_bytes, int64_t HW, int n, const float* u, void* ws, size_t ws_bytes,
                        int64_t* pix_out, int* labels_out, void* stream_) {
    if (HW < 0 || n < 0) return ISR_ERR_INVALID_ARG;
    if (n == 0) return ISR_OK;
    if (!labels || !u || !ws || !pix_out || !labels_out || HW == 0) return ISR_ERR_INVALID_ARG;
    if (ws_bytes < sampler_ws_bytes(HW)) return ISR_ERR_WORKSPACE;
    return launch_sampler(labels, label_bytes, HW, n, u, ws, pix_out, labels_out, static_cast<cudaStream_t>(stream_));
}

size_t isr_contrastive_workspace_bytes(int N, int F, int K) {
    if (N < 0 || F < 0 || K < 0) return 0;
    return contrastive_ws_bytes(N, F, K) + 256;
}

int isr_contrastive_forward(int N, int F, int K, const float* features, const int* labels, const float* predef_u,
                            float temp_lambda, int min_pixnum, void* ws, size_t ws_bytes, float* loss, void* stream_) {
    if (N < 0 || F <= 0 || K < 0 || !loss) return ISR_ERR_INVALID_ARG;
    if (F > ISR_MAX_EXTRA_DIMS) return ISR_ERR_UNSUPPORTED;
    if (N > 0 && (!features || !labels || !ws)) return ISR_ERR_INVALID_ARG;
    if (ws_bytes < contrastive_ws_bytes(N, F, K)) return ISR_ERR_WORKSPACE;
    return launch_contrastive_fwd(N, F, K, features, labels, predef_u, temp_lambda, min_pixnum, ws, loss,
                                  static_cast<cudaStream_t>(stream_));
}

int isr_contrastive_backward(int N, int F, int K, const float* features, const int* labels, const float* predef_u,
                             const void* ws, const float* grad_scale, float* dL_dfeatures, void* stream_) {
    (void)features;
    if (N < 0 || F <= 0 || K < 0) return ISR_ERR_INVALID_ARG;
    if (F > ISR_MAX_EXTRA_DIMS) return ISR_ERR_UNSUPPORTED;
    if (N > 0 && (!labels || !ws || !dL_dfeatures)) return ISR_ERR_INVALID_ARG;
    return launch_contrastive_bwd(N, F, K, labels, predef_u, ws, grad_scale, dL_dfeatures,
                                  static_cast<cudaStream_t>(stream_));
}

int isr_rownorm_forward(int P, int F, const float* x, float eps1, float eps2, int stages, float* y, void* stream_) {
    if (P < 0 || F < 0 || stages < 1 || stages > 2) return ISR_ERR_INVALID_ARG;
    if (F > ISR_MAX_EXTRA_DIMS) return ISR_ERR_UNSUPPORTED;
    if (P == 0 || F == 0) return ISR_OK;
    if (!x || !y) return ISR_ERR_INVALID_ARG;
    return launch_rownorm(true, P, F, x, nullptr, eps1, eps2, stages, y, static_cast<cudaStream_t>(stream_));
}

int isr_rownorm_backward(int P, int F, const float* x, const float* dy, float eps1, float eps2, int stages, float* dx,
                         void* stream_) {
    if (P < 0 || F < 0 || stages < 1 || stages > 2) return ISR_ERR_INVALID_ARG;
    if (F > ISR_MAX_EXTRA_DIMS) return ISR_ERR_UNSUPPORTED;
    if (P == 0 || F == 0) return ISR_OK;
    if (!x || !dy || !dx) return ISR_ERR_INVALID_ARG;
    return launch_rownorm(false, P, F, x, dy, eps1, eps2, stages, dx, static_cast<cudaStream_t>(stream_));
}

int isr_aux_maps_forward(int W, int H, const float* allmap, const float* normal_rot_host, const float* ray_mat_host,
                         float depth_ratio, float* rend_normal, float* rend_depth, float* rend_median, float* surf_depth,
                         float* surf_normal, void* stream_) {
    if (W <= 0 || H <= 0) return ISR_ERR_INVALID_ARG;
    if (!allmap || !normal_rot_host || !ray_mat_host || !rend_normal || !rend_depth || !rend_median || !surf_depth || !surf_normal)
        return ISR_ERR_INVALID_ARG;
    return launch_aux_fwd(W, H, allmap, normal_rot_host, ray_mat_host, depth_ratio, rend_normal, rend_depth, rend_median,
                          surf_depth, surf_normal, static_cast<cudaStream_t>(stream_));
}

int isr_aux_maps_backward(int W, int H, const float* allmap, const float* normal_rot_host, const float* ray_mat_host,
                          float depth_ratio, const float* g_rend_normal, const float* g_rend_depth,
                          const float* g_rend_median, const float* g_surf_depth, const float* g_surf_normal,
                          float* g_allmap, void* stream_) {
    if (W <= 0 || H <= 0) return ISR_ERR_INVALID_ARG;
    if (!allmap || !normal_rot_host || !ray_mat_host || !g_allmap) return ISR_ERR_INVALID_ARG;
    return launch_aux_bwd(W, H, allmap, normal_rot_host, ray_mat_host, depth_ratio, g_rend_normal, g_rend_depth,
                          g_rend_median, g_surf_depth, g_surf_normal, g_allmap, static_cast<cudaStream_t>(stream_));
}

int isr_adam_step(size_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                  float beta2, float eps, int step, const int* step_dev, void* stream_) {
    if (step < 1 && step_dev == nullptr) return ISR_ERR_INVALID_ARG;
    if (n == 0) return ISR_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq) return ISR_ERR_INVALID_ARG;
    if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) return ISR_ERR_INVALID_ARG;
    return launch_adam(n, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, step_dev,
                       static_cast<cudaStream_t>(stream_));
}

int isr_adam_rownorm_step(int P, int F, float* param, const float* dy, const float* grad_extra, float* exp_avg,
                          float* exp_avg_sq, float eps1, float eps2, int stages, float lr, float beta1, float beta2, float eps,
                          int step, const int* step_dev, void* stream_) {
    if (P < 0 || F < 0 || stages < 1 || stages > 2) return ISR_ERR_INVALID_ARG;
    if (step < 1 && step_dev == nullptr) return ISR_ERR_INVALID_ARG;
    if (F > ISR_MAX_EXTRA_DIMS) return ISR_ERR_UNSUPPORTED;
    if (P == 0 || F == 0) return ISR_OK;
    if (!param || !dy || !exp_avg || !exp_avg_sq) return ISR_ERR_INVALID_ARG;
    if ((F & 3) == 0 && (((uintptr_t)param | (uintptr_t)dy | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)grad_extra) & 15))
        return ISR_ERR_INVALID_ARG;
    return launch_adam_rownorm(P, F, param, dy, grad_extra, exp_avg, exp_avg_sq, eps1, eps2, stages, lr, beta1, beta2, eps, step,
                               step_dev, static_cast<cudaStream_t>(stream_));
}

size_t isr_knn_workspace_bytes(int P) { return P < 0 ? 0 : knn_ws_bytes(P) + 256; }

int isr_knn_mean_dist2(int P, const float* points, float* mean_dist2, void* ws, size_t ws_bytes, void* stream_) {
    if (P < 0) return ISR_ERR_INVALID_ARG;
    if (P == 0) return ISR_OK;
    if (!points || !mean_dist2 || !ws) return ISR_ERR_INVALID_ARG;
    if (ws_bytes < knn_ws_bytes(P)) return ISR_ERR_WORKSPACE;
    return launch_knn(P, points, mean_dist2, ws, ws_bytes, static_cast<cudaStream_t>(stream_));
}

size_t isr_photometric_workspace_bytes(int C, int H, int W) {
    return (C <= 0 || H <= 0 || W <= 0) ? 0 : photometric_ws_bytes(C, H, W);
}

int isr_photometric_forward(int C, int H, int W, const float* image, const float* gt, float lambda_dssim, void* ws,
                            size_t ws_bytes, float* out3, void* stream_) {
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535) return ISR_ERR_INVALID_ARG;
    if (!image || !gt || !ws || !out3) return ISR_ERR_INVALID_ARG;
    if (ws_bytes < photometric_ws_bytes(C, H, W)) return ISR_ERR_WORKSPACE;
    return launch_photometric_fwd(C, H, W, image, gt, lambda_dssim, ws, out3, static_cast<cudaStream_t>(stream_));
}

int isr_photometric_backward(int C, int H, int W, const float* image, const float* gt, float lambda_dssim, const void* ws,
                             const float* grad_scale, float* dL_dimage, void* stream_) {
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535) return ISR_ERR_INVALID_ARG;
    if (!image || !gt || !ws || !dL_dimage) return ISR_ERR_INVALID_ARG;
    return launch_photometric_bwd(C, H, W, image, gt, lambda_dssim, ws, grad_scale, dL_dimage,
                                  static_cast<cudaStream_t>(stream_));
}

int isr_densify_stats(int P, const int* radii, const float* dL_dmeans2D, float* max_radii2D, float* xyz_gradient_accum,
                      float* denom, void* stream_) {
    if (P < 0) return ISR_ERR_INVALID_ARG;
    if (P == 0) return ISR_OK;
    if (!radii || !dL_dmeans2D || !max_radii2D || !xyz_gradient_accum || !denom) return ISR_ERR_INVALID_ARG;
    return launch_densify_stats(P, radii, dL_dmeans2D, max_radii2D, xyz_gradient_accum, denom, static_cast<cudaStream_t>(stream_));
}

size_t isr_tracker_workspace_bytes(int P, int K) { return (P < 0 || K < 0) ? 0 : tracker_ws_bytes(P, K) + 256; }

int isr_tracker_mark(const int* pairs, int64_t n_pairs, const int* seg_rows, int64_t HW, int P, int K, void* ws,
                     size_t ws_bytes, int* counts, void* stream_) {
    if (P < 0 || K < 1 || n_pairs < 0 || HW < 0) return ISR_ERR_INVALID_ARG;
    if (!ws || !counts || (n_pairs > 0 && (!pairs || !seg_rows))) return ISR_ERR_INVALID_ARG;
    if (ws_bytes < tracker_ws_bytes(P, K)) return ISR_ERR_WORKSPACE;
    return launch_tracker_mark(pairs, n_pairs, seg_rows, HW, P, K, ws, counts, static_cast<cudaStream_t>(stream_));
}

int isr_tracker_fill(int P, int K, const void* ws, const int64_t* row_offsets, int* out_ids, void* stream_) {
    if (P < 0 || K < 1) return ISR_ERR_INVALID_ARG;
    if (!ws || !row_offsets) return ISR_ERR_INVALID_ARG;
    if (P == 0) return ISR_OK;
    return launch_tracker_fill(P, K, ws, row_offsets, out_ids, static_cast<cudaStream_t>(stream_));
}

}  // extern "C"
