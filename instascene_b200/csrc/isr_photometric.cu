// isr_photometric.cu -- fused photometric loss of the RGB training step:
//     loss = (1 - lambda) * mean|I - G|  +  lambda * (1 - SSIM(I, G))            train.py:76-77
// with SSIM exactly as utils/loss_utils.py:39-83: 11x11 Gaussian window (sigma 1.5, zero padding), per channel,
// C1 = 0.01^2, C2 = 0.03^2, mean over all C*H*W positions.
//
// The reference runs 5 depthwise conv2d + ~20 elementwise kernels forward and their autograd mirror backward
// (~1.2 GB of HBM traffic at 1080p).  Here: ONE forward kernel (reads I, G once; separable 11-tap window through
// shared memory; writes three derivative maps + per-block partial sums) and ONE backward kernel (reads the three maps,
// I and G once; writes dL/dI).  HBM bound: algorithmic bytes fwd = CHW*4*(2 + 3), bwd = CHW*4*(3 + 2 + 1).
// The loss is reduced in a fixed order (per-block partials, then one block) -> bitwise reproducible run to run.
#include "isr_common.cuh"

namespace isr {

namespace {
constexpr int kWin = 11, kHalo = 5;
constexpr int kBX = 32, kBY = 16;                       // pixels per CTA (one thread per pixel, 512 threads)
constexpr int kSX = kBX + 2 * kHalo, kSY = kBY + 2 * kHalo;  // staged tile incl. halo: 42 x 26
// gaussian(11, 1.5) of utils/loss_utils.py:31-33 evaluated like the reference (float32 tensor / its float32 sum)
__constant__ float c_win[kWin] = {1.028380124e-03f, 7.598758209e-03f, 3.600077331e-02f, 1.093606874e-01f,
                                  2.130055279e-01f, 2.660117149e-01f, 2.130055279e-01f, 1.093606874e-01f,
                                  3.600077331e-02f, 7.598758209e-03f, 1.028380124e-03f};
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float s = 0.0f;
    if (tid < 32) {
        const int nw = (blockDim.x * blockDim.y) >> 5;
        s = tid < nw ? red[tid] : 0.0f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    }
    return s;  // valid on thread 0
}
}  // namespace

// maps[3][C*H*W]: d ssim_map / d mu1 (total, through sigma1_sq and sigma12 too), d / d E[I*I], d / d E[I*G]
__global__ void __launch_bounds__(kBX * kBY)
photometric_fwd_kernel(int C, int H, int W, const float* __restrict__ img, const float* __restrict__ gt,
                       float* __restrict__ maps, float* __restrict__ partials /*[blocks][2]: sum|I-G|, sum ssim*/) {
    __shared__ float sI[kSY][kSX + 1], sG[kSY][kSX + 1];
    __shared__ float hsum[5][kSY][kBX + 1];
    __shared__ float red[32];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * kBX, y0 = blockIdx.y * kBY;
    const size_t plane = (size_t)H * W;
    const float* ip = img + (size_t)c * plane;
    const float* gp = gt + (size_t)c * plane;
    const int tid = threadIdx.y * kBX + threadIdx.x;
    for (int i = tid; i < kSX * kSY; i += kBX * kBY) {
        const int sy = i / kSX, sx = i % kSX;
        const int gx = x0 + sx - kHalo, gy = y0 + sy - kHalo;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;  // zero padding (F.conv2d padding = 5)
        sI[sy][sx] = in ? __ldg(ip + (size_t)gy * W + gx) : 0.0f;
        sG[sy][sx] = in ? __ldg(gp + (size_t)gy * W + gx) : 0.0f;
    }
    __syncthreads();
    // horizontal pass: kSY rows x kBX columns
    for (int i = tid; i < kSY * kBX; i += kBX * kBY) {
        const int sy = i / kBX, sx = i % kBX;
        float a = 0, b = 0, aa = 0, bb = 0, ab = 0;
#pragma unroll
        for (int k = 0; k < kWin; k++) {
            const float w = c_win[k], p = sI[sy][sx + k], q = sG[sy][sx + k];
            a = fmaf(w, p, a); b = fmaf(w, q, b);
            aa = fmaf(w, p * p, aa); bb = fmaf(w, q * q, bb); ab = fmaf(w, p * q, ab);
        }
        hsum[0][sy][sx] = a; hsum[1][sy][sx] = b; hsum[2][sy][sx] = aa; hsum[3][sy][sx] = bb; hsum[4][sy][sx] = ab;
    }
    __syncthreads();
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int gx = x0 + tx, gy = y0 + ty;
    float l1 = 0.0f, ss = 0.0f;
    if (gx < W && gy < H) {
        float mu1 = 0, mu2 = 0, e11 = 0, e22 = 0, e12 = 0;
#pragma unroll
        for (int k = 0; k < kWin; k++) {
            const float w = c_win[k];
            mu1 = fmaf(w, hsum[0][ty + k][tx], mu1); mu2 = fmaf(w, hsum[1][ty + k][tx], mu2);
            e11 = fmaf(w, hsum[2][ty + k][tx], e11); e22 = fmaf(w, hsum[3][ty + k][tx], e22);
            e12 = fmaf(w, hsum[4][ty + k][tx], e12);
        }
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu12;
        const float A = 2.0f * mu12 + kC1, B = 2.0f * s12 + kC2, Cc = mu1_sq + mu2_sq + kC1, D = s1 + s2 + kC2;
        const float inv = 1.0f / (Cc * D);
        const float m = A * B * inv;
        ss = m;
        const float p = sI[ty + kHalo][tx + kHalo], q = sG[ty + kHalo][tx + kHalo];
        l1 = fabsf(p - q);
        // partial derivatives of m w.r.t. (mu1, E[II], E[IG]) with sigma1_sq = E[II] - mu1^2, sigma12 = E[IG] - mu1*mu2
        const float dm_ds1 = -m / D;
        const float dm_ds12 = 2.0f * A * inv;
        const float dm_dmu1 = 2.0f * mu2 * B * inv - 2.0f * mu1 * m / Cc - 2.0f * mu1 * dm_ds1 - mu2 * dm_ds12;
        const size_t o = (size_t)c * plane + (size_t)gy * W + gx;
        const size_t chw = (size_t)C * plane;
        maps[o] = dm_dmu1;
        maps[o + chw] = dm_ds1;
        maps[o + 2 * chw] = dm_ds12;
    }
    const float bl1 = block_sum(l1, red);
    const float bss = block_sum(ss, red);
    if (tid == 0) {
        const size_t b = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partials[2 * b] = bl1;
        partials[2 * b + 1] = bss;
    }
}

// out[0] = loss, out[1] = L1 mean, out[2] = SSIM mean; fixed summation order (one block, strided then tree)
__global__ void __launch_bounds__(1024)
photometric_reduce_kernel(const float* __restrict__ partials, int nblocks, float inv_n, float lambda, float* __restrict__ out) {
    __shared__ double r1[1024], r2[1024];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 1024) { a += partials[2 * i]; b += partials[2 * i + 1]; }
    r1[threadIdx.x] = a; r2[threadIdx.x] = b;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) { r1[threadIdx.x] += r1[threadIdx.x + s]; r2[threadIdx.x] += r2[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float l1 = (float)(r1[0] * (double)inv_n), ssim = (float)(r2[0] * (double)inv_n);
        out[0] = (1.0f - lambda) * l1 + lambda * (1.0f - ssim);
        out[1] = l1;
        out[2] = ssim;
    }
}

// dL/dI = g * [ (1-lambda)/N * sign(I-G)  -  lambda/N * ( W*M1 + 2 I (W*M2) + G (W*M3) ) ],  W* = the same window
__global__ void __launch_bounds__(kBX * kBY)
photometric_bwd_kernel(int C, int H, int W, const float* __restrict__ img, const float* __restrict__ gt,
                       const float* __restrict__ maps, const float* __restrict__ grad_scale, float lambda, float inv_n,
                       float* __restrict__ dimg) {
    __shared__ float sM[3][kSY][kSX + 1];
    __shared__ float hsum[3][kSY][kBX + 1];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * kBX, y0 = blockIdx.y * kBY;
    const size_t plane = (size_t)H * W, chw = (size_t)C * plane;
    const float* mp = maps + (size_t)c * plane;
    const int tid = threadIdx.y * kBX + threadIdx.x;
    for (int i = tid; i < kSX * kSY; i += kBX * kBY) {
        const int sy = i / kSX, sx = i % kSX;
        const int gx = x0 + sx - kHalo, gy = y0 + sy - kHalo;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        const size_t o = (size_t)gy * W + gx;
        sM[0][sy][sx] = in ? __ldg(mp + o) : 0.0f;
        sM[1][sy][sx] = in ? __ldg(mp + chw + o) : 0.0f;
        sM[2][sy][sx] = in ? __ldg(mp + 2 * chw + o) : 0.0f;
    }
    __syncthreads();
    for (int i = tid; i < kSY * kBX; i += kBX * kBY) {
        const int sy = i / kBX, sx = i % kBX;
        float a = 0, b = 0, d = 0;
#pragma unroll
        for (int k = 0; k < kWin; k++) {
            const float w = c_win[k];
            a = fmaf(w, sM[0][sy][sx + k], a); b = fmaf(w, sM[1][sy][sx + k], b); d = fmaf(w, sM[2][sy][sx + k], d);
        }
        hsum[0][sy][sx] = a; hsum[1][sy][sx] = b; hsum[2][sy][sx] = d;
    }
    __syncthreads();
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx < W && gy < H) {
        float a = 0, b = 0, d = 0;
#pragma unroll
        for (int k = 0; k < kWin; k++) {
            const float w = c_win[k];
            a = fmaf(w, hsum[0][ty + k][tx], a); b = fmaf(w, hsum[1][ty + k][tx], b); d = fmaf(w, hsum[2][ty + k][tx], d);
        }
        const size_t o = (size_t)c * plane + (size_t)gy * W + gx;
        const float p = __ldg(img + o), q = __ldg(gt + o);
        const float diff = p - q;
        const float sgn = diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f);  // torch.abs' subgradient: sign(0) = 0
        const float g = grad_scale ? __ldg(grad_scale) : 1.0f;
        dimg[o] = g * inv_n * ((1.0f - lambda) * sgn - lambda * (a + 2.0f * p * b + q * d));
    }
}

// Densification statistics of the RGB training step (train.py:139-142, scene/gaussian_model.py:602-605) in one pass:
// for visible Gaussians (radii > 0): max_radii2D = max(max_radii2D, radii); xyz_gradient_accum += |dL/dmeans2D|;
// denom += 1.  Three masked gathers/scatters + a norm in the reference (each a boolean-index kernel chain with a
// host sync for the mask count).
__global__ void __launch_bounds__(256)
densify_stats_kernel(int P, const int* __restrict__ radii, const float* __restrict__ grad2d, float* __restrict__ max_radii2D,
                     float* __restrict__ grad_accum, float* __restrict__ denom) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;
    max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
    const float gx = grad2d[3 * (size_t)i], gy = grad2d[3 * (size_t)i + 1], gz = grad2d[3 * (size_t)i + 2];
    grad_accum[i] += sqrtf(gx * gx + gy * gy + gz * gz);
    denom[i] += 1.0f;
}

int launch_densify_stats(int P, const int* radii, const float* grad2d, float* max_radii2D, float* grad_accum, float* denom,
                         cudaStream_t stream) {
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, radii, grad2d, max_radii2D, grad_accum, denom);
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

size_t photometric_ws_bytes(int C, int H, int W) {
    const size_t chw = (size_t)C * H * W;
    const size_t blocks = (size_t)((W + kBX - 1) / kBX) * ((H + kBY - 1) / kBY) * C;
    return align_up(3 * chw * 4, 256) + align_up(blocks * 8, 256);
}

int launch_photometric_fwd(int C, int H, int W, const float* img, const float* gt, float lambda, void* ws, float* out,
                           cudaStream_t stream) {
    const size_t chw = (size_t)C * H * W;
    float* maps = static_cast<float*>(ws);
    float* partials = reinterpret_cast<float*>(static_cast<char*>(ws) + align_up(3 * chw * 4, 256));
    const dim3 grid((W + kBX - 1) / kBX, (H + kBY - 1) / kBY, C), block(kBX, kBY);
    photometric_fwd_kernel<<<grid, block, 0, stream>>>(C, H, W, img, gt, maps, partials);
    photometric_reduce_kernel<<<1, 1024, 0, stream>>>(partials, (int)(grid.x * grid.y * grid.z), 1.0f / (float)chw, lambda, out);
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_photometric_bwd(int C, int H, int W, const float* img, const float* gt, float lambda, const void* ws,
                           const float* grad_scale, float* dimg, cudaStream_t stream) {
    const size_t chw = (size_t)C * H * W;
    const dim3 grid((W + kBX - 1) / kBX, (H + kBY - 1) / kBY, C), block(kBX, kBY);
    photometric_bwd_kernel<<<grid, block, 0, stream>>>(C, H, W, img, gt, static_cast<const float*>(ws), grad_scale, lambda,
                                                       1.0f / (float)chw, dimg);
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

}  // namespace isr
