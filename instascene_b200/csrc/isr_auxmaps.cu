// isr_auxmaps.cu -- fused post-processing of the rasterizer's 7-channel `allmap` into the derived maps that render()
// returns (gaussian_renderer/__init__.py:127-156 + utils/point_utils.py:10-40), forward and backward.
// The reference does this with ~40 torch kernels per view (permute/matmul for the normal rotation, two nan_to_num,
// a meshgrid + two 3x3 inverses + matmuls for the pixel rays, slicing/cross/normalize for the stencil) and as many
// again in autograd; here it is one kernel each way, one thread per pixel.
//
//   rend_normal  = normal_view @ M                        M = world_view_transform[:3,:3].T  (view -> world)
//   rend_median  = nan_to_num(allmap[5], nan=0, posinf=0)
//   rend_depth   = nan_to_num(allmap[0] / allmap[1], nan=0, posinf=0)            (expected depth)
//   surf_depth   = (1 - r) * rend_depth + r * rend_median                        r = pipe.depth_ratio
//   surf_normal  = normalize(cross(P[y+1,x] - P[y-1,x], P[y,x+1] - P[y,x-1])) * alpha.detach()   (0 on the border)
//                  with P[y,x] = surf_depth[y,x] * ([x, y, 1] @ K) + origin     (origin cancels in the differences)
#include "isr_common.cuh"

namespace isr {

struct AuxConsts {
    float M[9];  // row-major: n_world[c] = sum_k n_view[k] * M[3k + c]
    float K[9];  // row-major: ray[c]     = x * K[c] + y * K[3 + c] + K[6 + c]
    float depth_ratio;
};

__device__ __forceinline__ float nan_posinf_to_zero(float v) {
    // torch.nan_to_num(v, 0, 0): NaN -> 0, +inf -> 0, -inf -> lowest finite float
    if (v != v) return 0.0f;
    if (v == __int_as_float(0x7f800000)) return 0.0f;
    if (v == __int_as_float(0xff800000)) return -3.4028234663852886e38f;
    return v;
}

__device__ __forceinline__ float surf_depth_at(const float* __restrict__ allmap, size_t HW, size_t pix, float r) {
    const float D = __ldg(allmap + pix), A = __ldg(allmap + HW + pix), med = __ldg(allmap + 5 * HW + pix);
    return nan_posinf_to_zero(D / A) * (1.0f - r) + r * nan_posinf_to_zero(med);
}

__device__ __forceinline__ void ray_at(const AuxConsts& c, int x, int y, float* rd) {
    const float fx = (float)x, fy = (float)y;
#pragma unroll
    for (int k = 0; k < 3; k++) rd[k] = fx * c.K[k] + fy * c.K[3 + k] + c.K[6 + k];
}

// a = P(down) - P(up), b = P(right) - P(left) of the centre pixel (cx, cy); sd values passed in
__device__ __forceinline__ void stencil_vectors(const AuxConsts& c, int cx, int cy, float sd_up, float sd_down,
                                                float sd_left, float sd_right, float* a, float* b) {
    float ru[3], rdn[3], rl[3], rr[3];
    ray_at(c, cx, cy - 1, ru); ray_at(c, cx, cy + 1, rdn); ray_at(c, cx - 1, cy, rl); ray_at(c, cx + 1, cy, rr);
#pragma unroll
    for (int k = 0; k < 3; k++) { a[k] = sd_down * rdn[k] - sd_up * ru[k]; b[k] = sd_right * rr[k] - sd_left * rl[k]; }
}

__global__ void __launch_bounds__(256)
aux_maps_fwd_kernel(int W, int H, const float* __restrict__ allmap, const AuxConsts c, float* __restrict__ rend_normal,
                    float* __restrict__ rend_depth, float* __restrict__ rend_median, float* __restrict__ surf_depth,
                    float* __restrict__ surf_normal) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)W * H, pix = (size_t)y * W + x;
    const float D = __ldg(allmap + pix), A = __ldg(allmap + HW + pix);
    const float nv0 = __ldg(allmap + 2 * HW + pix), nv1 = __ldg(allmap + 3 * HW + pix), nv2 = __ldg(allmap + 4 * HW + pix);
#pragma unroll
    for (int ch = 0; ch < 3; ch++) rend_normal[ch * HW + pix] = nv0 * c.M[ch] + nv1 * c.M[3 + ch] + nv2 * c.M[6 + ch];
    const float med = nan_posinf_to_zero(__ldg(allmap + 5 * HW + pix));
    const float expd = nan_posinf_to_zero(D / A);
    rend_median[pix] = med;
    rend_depth[pix] = expd;
    surf_depth[pix] = expd * (1.0f - c.depth_ratio) + c.depth_ratio * med;
    float n[3] = {0.0f, 0.0f, 0.0f};
    if (x >= 1 && x <= W - 2 && y >= 1 && y <= H - 2) {
        float a[3], b[3];
        stencil_vectors(c, x, y, surf_depth_at(allmap, HW, pix - W, c.depth_ratio), surf_depth_at(allmap, HW, pix + W, c.depth_ratio),
                        surf_depth_at(allmap, HW, pix - 1, c.depth_ratio), surf_depth_at(allmap, HW, pix + 1, c.depth_ratio), a, b);
        const float cr[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        const float len = fmaxf(sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]), 1e-12f);  // F.normalize eps
#pragma unroll
        for (int k = 0; k < 3; k++) n[k] = cr[k] / len * A;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) surf_normal[k * HW + pix] = n[k];
}

// gradient that the normal at centre (cx,cy) sends to the stencil vectors a and b
__device__ __forceinline__ bool centre_grads(const AuxConsts& c, const float* __restrict__ allmap,
                                             const float* __restrict__ g_surf_normal, int W, int H, int cx, int cy,
                                             float* ga, float* gb) {
    if (cx < 1 || cx > W - 2 || cy < 1 || cy > H - 2) return false;
    const size_t HW = (size_t)W * H, cp = (size_t)cy * W + cx;
    float a[3], b[3];
    stencil_vectors(c, cx, cy, surf_depth_at(allmap, HW, cp - W, c.depth_ratio), surf_depth_at(allmap, HW, cp + W, c.depth_ratio),
                    surf_depth_at(allmap, HW, cp - 1, c.depth_ratio), surf_depth_at(allmap, HW, cp + 1, c.depth_ratio), a, b);
    const float alpha = __ldg(allmap + HW + cp);
    const float G[3] = {__ldg(g_surf_normal + cp) * alpha, __ldg(g_surf_normal + HW + cp) * alpha,
                        __ldg(g_surf_normal + 2 * HW + cp) * alpha};
    const float cr[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const float nrm = sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
    float gn[3];
    if (nrm > 1e-12f) {  // d(v/|v|) = (I - n n^T)/|v|
        const float inv = 1.0f / nrm;
        const float nn[3] = {cr[0] * inv, cr[1] * inv, cr[2] * inv};
        const float dot = nn[0] * G[0] + nn[1] * G[1] + nn[2] * G[2];
#pragma unroll
        for (int k = 0; k < 3; k++) gn[k] = (G[k] - nn[k] * dot) * inv;
    } else {  // clamped denominator: v / 1e-12
#pragma unroll
        for (int k = 0; k < 3; k++) gn[k] = G[k] * 1e12f;
    }
    // n = a x b:  g_a = b x g_n,  g_b = g_n x a
    ga[0] = b[1] * gn[2] - b[2] * gn[1]; ga[1] = b[2] * gn[0] - b[0] * gn[2]; ga[2] = b[0] * gn[1] - b[1] * gn[0];
    gb[0] = gn[1] * a[2] - gn[2] * a[1]; gb[1] = gn[2] * a[0] - gn[0] * a[2]; gb[2] = gn[0] * a[1] - gn[1] * a[0];
    return true;
}

__global__ void __launch_bounds__(256)
aux_maps_bwd_kernel(int W, int H, const float* __restrict__ allmap, const AuxConsts c,
                    const float* __restrict__ g_rend_normal, const float* __restrict__ g_rend_depth,
                    const float* __restrict__ g_rend_median, const float* __restrict__ g_surf_depth,
                    const float* __restrict__ g_surf_normal, float* __restrict__ g_allmap) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)W * H, pix = (size_t)y * W + x;
    float g_sd = g_surf_depth ? __ldg(g_surf_depth + pix) : 0.0f;
    if (g_surf_normal) {
        float gp[3] = {0.0f, 0.0f, 0.0f}, ga[3], gb[3];
        // this pixel is the DOWN neighbour of (x, y-1), the UP neighbour of (x, y+1),
        // the RIGHT neighbour of (x-1, y) and the LEFT neighbour of (x+1, y)
        if (centre_grads(c, allmap, g_surf_normal, W, H, x, y - 1, ga, gb)) { gp[0] += ga[0]; gp[1] += ga[1]; gp[2] += ga[2]; }
        if (centre_grads(c, allmap, g_surf_normal, W, H, x, y + 1, ga, gb)) { gp[0] -= ga[0]; gp[1] -= ga[1]; gp[2] -= ga[2]; }
        if (centre_grads(c, allmap, g_surf_normal, W, H, x - 1, y, ga, gb)) { gp[0] += gb[0]; gp[1] += gb[1]; gp[2] += gb[2]; }
        if (centre_grads(c, allmap, g_surf_normal, W, H, x + 1, y, ga, gb)) { gp[0] -= gb[0]; gp[1] -= gb[1]; gp[2] -= gb[2]; }
        float rd[3];
        ray_at(c, x, y, rd);
        g_sd += gp[0] * rd[0] + gp[1] * rd[1] + gp[2] * rd[2];
    }
    const float g_exp = (1.0f - c.depth_ratio) * g_sd + (g_rend_depth ? __ldg(g_rend_depth + pix) : 0.0f);
    const float g_med = c.depth_ratio * g_sd + (g_rend_median ? __ldg(g_rend_median + pix) : 0.0f);
    const float D = __ldg(allmap + pix), A = __ldg(allmap + HW + pix), med_raw = __ldg(allmap + 5 * HW + pix);
    const float e_raw = D / A;
    const bool e_fin = isfinite(e_raw);
    // torch: d nan_to_num = grad * isfinite(input); d(D/A) = (g/A, -g*D/A^2).  Where the quotient is not finite torch
    // produces 0/0 = NaN for the A = 0 pixels; nothing contributes there (no Gaussian reached the pixel), 0 is written.
    g_allmap[pix] = e_fin ? g_exp / A : 0.0f;
    g_allmap[HW + pix] = e_fin ? -g_exp * D / (A * A) : 0.0f;
    float gn[3] = {0.0f, 0.0f, 0.0f};
    if (g_rend_normal) {
        const float g0 = __ldg(g_rend_normal + pix), g1 = __ldg(g_rend_normal + HW + pix), g2 = __ldg(g_rend_normal + 2 * HW + pix);
#pragma unroll
        for (int k = 0; k < 3; k++) gn[k] = g0 * c.M[3 * k] + g1 * c.M[3 * k + 1] + g2 * c.M[3 * k + 2];
    }
    g_allmap[2 * HW + pix] = gn[0];
    g_allmap[3 * HW + pix] = gn[1];
    g_allmap[4 * HW + pix] = gn[2];
    g_allmap[5 * HW + pix] = isfinite(med_raw) ? g_med : 0.0f;
    g_allmap[6 * HW + pix] = 0.0f;
}

static AuxConsts make_consts(const float* M_host, const float* K_host, float depth_ratio) {
    AuxConsts c;
    for (int i = 0; i < 9; i++) { c.M[i] = M_host[i]; c.K[i] = K_host[i]; }
    c.depth_ratio = depth_ratio;
    return c;
}

int launch_aux_fwd(int W, int H, const float* allmap, const float* M_host, const float* K_host, float depth_ratio,
                   float* rend_normal, float* rend_depth, float* rend_median, float* surf_depth, float* surf_normal,
                   cudaStream_t stream) {
    const dim3 grid((W + 255) / 256, H);
    aux_maps_fwd_kernel<<<grid, 256, 0, stream>>>(W, H, allmap, make_consts(M_host, K_host, depth_ratio), rend_normal,
                                                  rend_depth, rend_median, surf_depth, surf_normal); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_aux_bwd(int W, int H, const float* allmap, const float* M_host, const float* K_host, float depth_ratio,
                   const float* g_rend_normal, const float* g_rend_depth, const float* g_rend_median,
                   const float* g_surf_depth, const float* g_surf_normal, float* g_allmap, cudaStream_t stream) {
    const dim3 grid((W + 255) / 256, H);
    aux_maps_bwd_kernel<<<grid, 256, 0, stream>>>(W, H, allmap, make_consts(M_host, K_host, depth_ratio), g_rend_normal,
                                                  g_rend_depth, g_rend_median, g_surf_depth, g_surf_normal, g_allmap); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

}  // namespace isr
