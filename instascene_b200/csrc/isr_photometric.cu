// isr_photometric.cu -- fused photometric loss of the RGB training step:
//     loss = (1 - lambda) * mean|I - G|  +  lambda * (1 - SSIM(I, G))            train.py:76-77
// with SSIM exactly as utils/loss_utils.py:39-83: 11x11 Gaussian window (sigma 1.5, zero padding), per channel,
// C1 = 0.01^2, C2 = 0.03^2, mean over all C*H*W positions.
//
// The reference runs 5 depthwise conv2d + ~20 elementwise kernels forward and their autograd mirror backward
// (~1.2 GB of HBM traffic at 1080p).  Here: ONE forward kernel (reads I, G once; separable 11-tap window through
// shared memory; writes three derivative maps + per-block partial sums) and ONE backward kernel (reads the three maps,
// I and G once; writes dL/dI).  Algorithmic bytes fwd = CHW*4*(2 + 3), bwd = CHW*4*(3 + 2 + 1); with ~150 FMA per
// pixel for the five windowed moments the forward is instruction-issue bound at ~20 % of the HBM roofline (92 us at
// 1080p), the backward reaches 33 % (68 us).
// The loss is reduced in a fixed order (per-block partials, then one block) -> bitwise reproducible run to run.
#include "isr_common.cuh"

namespace isr {

namespace {
constexpr int kWin = 11, kHalo = 5;
constexpr int kBX = 32;                                      // output pixels per CTA: 32 x BY (BY = 16 or 32)
constexpr int kSX = kBX + 2 * kHalo;                         // staged tile incl. halo: 42 x (BY + 10)
constexpr int kHO = 8;                                       // outputs per thread in the horizontal pass (18 inputs)
constexpr int kVO = 4;                                       // outputs per thread in the vertical pass (14 inputs)
template <int BY>
struct Tile {
    static constexpr int kBY = BY, kSY = BY + 2 * kHalo, kThreads = kBX * (BY / kVO);  // one (column, 4-row group) per thread
    static_assert(kSY * (kBX / kHO) <= kThreads, "horizontal pass: one (row, 8-column segment) per thread");
};
// gaussian(11, 1.5) of utils/loss_utils.py:31-33 evaluated like the reference (float32 tensor / its float32 sum)
__constant__ float c_win[kWin] = {1.028380124e-03f, 7.598758209e-03f, 3.600077331e-02f, 1.093606874e-01f,
                                  2.130055279e-01f, 2.660117149e-01f, 2.130055279e-01f, 1.093606874e-01f,
                                  3.600077331e-02f, 7.598758209e-03f, 1.028380124e-03f};
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

template <int kThreads>
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int tid = threadIdx.x;
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float s = 0.0f;
    if (tid < 32) {
        s = tid < kThreads / 32 ? red[tid] : 0.0f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    }
    return s;  // valid on thread 0
}

// zero-padded (F.conv2d padding = 5) copy of the 42 x (BY+10) neighbourhood of one plane's tile into shared memory.
// A thread keeps ONE column and walks down the rows (kThreads / 42 rows per step): no per-element division, and the
// copies are 4-byte cp.async (LDGSTS) issued back to back -- the first version loaded through a register and stored,
// one dependent global-load latency per element (46 % of the kernel's stall samples sat on that store).
// The caller waits (cp.async.wait_all + __syncthreads) once after staging all planes.
template <int kSY, int kThreads>
__device__ __forceinline__ void stage_tile(const float* __restrict__ plane, int x0, int y0, int W, int H,
                                           float (*dst)[kSX + 1]) {
    constexpr int kRowsPerStep = kThreads / kSX;
    const int sx = threadIdx.x % kSX, r = threadIdx.x / kSX;
    if (r >= kRowsPerStep) return;
    const int gx = x0 + sx - kHalo;
    const bool xin = gx >= 0 && gx < W;
#pragma unroll
    for (int sy0 = 0; sy0 < kSY; sy0 += kRowsPerStep) {
        const int sy = sy0 + r, gy = y0 + sy - kHalo;
        if (sy < kSY) {
            if (xin && gy >= 0 && gy < H) {
                const unsigned d = (unsigned)__cvta_generic_to_shared(&dst[sy][sx]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(plane + (size_t)gy * W + gx));
            } else {
                dst[sy][sx] = 0.0f;
            }
        }
    }
}
__device__ __forceinline__ void stage_wait() {
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();
}
}  // namespace

// Register-blocked separable window: in the horizontal pass a thread owns 8 consecutive outputs of one staged row (18
// inputs from shared memory instead of 88), in the vertical pass 4 consecutive outputs of one column (14 instead of 44);
// rows are padded to an odd stride so that both access patterns are bank-conflict free.
// maps[3][C*H*W]: d ssim_map / d mu1 (total, through sigma1_sq and sigma12 too), d / d E[I*I], d / d E[I*G]
template <int BY>
__global__ void __launch_bounds__(Tile<BY>::kThreads, 512 / Tile<BY>::kThreads)
photometric_fwd_kernel(int C, int H, int W, const float* __restrict__ img, const float* __restrict__ gt,
                       float* __restrict__ maps, float* __restrict__ partials /*[blocks][2]: sum|I-G|, sum ssim*/) {
    constexpr int kBY = Tile<BY>::kBY, kSY = Tile<BY>::kSY, kThreads = Tile<BY>::kThreads;
    __shared__ float sI[kSY][kSX + 1], sG[kSY][kSX + 1];
    __shared__ float hsum[5][kSY][kBX + 1];
    __shared__ float red[32];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * kBX, y0 = blockIdx.y * kBY;
    const size_t plane = (size_t)H * W;
    const int tid = threadIdx.x;
    stage_tile<kSY, kThreads>(img + (size_t)c * plane, x0, y0, W, H, sI);
    stage_tile<kSY, kThreads>(gt + (size_t)c * plane, x0, y0, W, H, sG);
    stage_wait();
    if (tid < kSY * (kBX / kHO)) {  // horizontal pass
        const int row = tid / (kBX / kHO), c0 = (tid % (kBX / kHO)) * kHO;
        float p[kHO + kWin - 1], q[kHO + kWin - 1];
#pragma unroll
        for (int j = 0; j < kHO + kWin - 1; j++) { p[j] = sI[row][c0 + j]; q[j] = sG[row][c0 + j]; }
#pragma unroll
        for (int o = 0; o < kHO; o++) {
            float a = 0, b = 0, aa = 0, bb = 0, ab = 0;
#pragma unroll
            for (int k = 0; k < kWin; k++) {
                const float w = c_win[k], pv = p[o + k], qv = q[o + k];
                a = fmaf(w, pv, a); b = fmaf(w, qv, b);
                aa = fmaf(w, pv * pv, aa); bb = fmaf(w, qv * qv, bb); ab = fmaf(w, pv * qv, ab);
            }
            hsum[0][row][c0 + o] = a; hsum[1][row][c0 + o] = b; hsum[2][row][c0 + o] = aa;
            hsum[3][row][c0 + o] = bb; hsum[4][row][c0 + o] = ab;
        }
    }
    __syncthreads();
    // vertical pass + SSIM map
    const int tx = tid % kBX, r0 = (tid / kBX) * kVO;
    float mu1[kVO], mu2[kVO], e11[kVO], e22[kVO], e12[kVO];
#pragma unroll
    for (int o = 0; o < kVO; o++) { mu1[o] = 0; mu2[o] = 0; e11[o] = 0; e22[o] = 0; e12[o] = 0; }
#pragma unroll
    for (int j = 0; j < kVO + kWin - 1; j++) {
        const float h0 = hsum[0][r0 + j][tx], h1 = hsum[1][r0 + j][tx], h2 = hsum[2][r0 + j][tx], h3 = hsum[3][r0 + j][tx],
                    h4 = hsum[4][r0 + j][tx];
#pragma unroll
        for (int o = 0; o < kVO; o++) {
            const int k = j - o;
            if (k >= 0 && k < kWin) {
                const float w = c_win[k];
                mu1[o] = fmaf(w, h0, mu1[o]); mu2[o] = fmaf(w, h1, mu2[o]); e11[o] = fmaf(w, h2, e11[o]);
                e22[o] = fmaf(w, h3, e22[o]); e12[o] = fmaf(w, h4, e12[o]);
            }
        }
    }
    float l1 = 0.0f, ss = 0.0f;
    const int gx = x0 + tx;
    const size_t chw = (size_t)C * plane;
#pragma unroll
    for (int o = 0; o < kVO; o++) {
        const int gy = y0 + r0 + o;
        if (gx < W && gy < H) {
            const float mu1_sq = mu1[o] * mu1[o], mu2_sq = mu2[o] * mu2[o], mu12 = mu1[o] * mu2[o];
            const float s1 = e11[o] - mu1_sq, s2 = e22[o] - mu2_sq, s12 = e12[o] - mu12;
            const float A = 2.0f * mu12 + kC1, B = 2.0f * s12 + kC2, Cc = mu1_sq + mu2_sq + kC1, D = s1 + s2 + kC2;
            const float inv = 1.0f / (Cc * D);
            const float m = A * B * inv;
            ss += m;
            l1 += fabsf(sI[r0 + o + kHalo][tx + kHalo] - sG[r0 + o + kHalo][tx + kHalo]);
            // partial derivatives of m w.r.t. (mu1, E[II], E[IG]) with sigma1_sq = E[II] - mu1^2, sigma12 = E[IG] - mu1*mu2
            const float dm_ds1 = -m / D;
            const float dm_ds12 = 2.0f * A * inv;
            const float dm_dmu1 = 2.0f * mu2[o] * B * inv - 2.0f * mu1[o] * m / Cc - 2.0f * mu1[o] * dm_ds1 - mu2[o] * dm_ds12;
            const size_t off = (size_t)c * plane + (size_t)gy * W + gx;
            maps[off] = dm_dmu1;
            maps[off + chw] = dm_ds1;
            maps[off + 2 * chw] = dm_ds12;
        }
    }
    const float bl1 = block_sum<kThreads>(l1, red);
    const float bss = block_sum<kThreads>(ss, red);
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
template <int BY>
__global__ void __launch_bounds__(Tile<BY>::kThreads, 512 / Tile<BY>::kThreads)
photometric_bwd_kernel(int C, int H, int W, const float* __restrict__ img, const float* __restrict__ gt,
                       const float* __restrict__ maps, const float* __restrict__ grad_scale, float lambda, float inv_n,
                       float* __restrict__ dimg) {
    constexpr int kBY = Tile<BY>::kBY, kSY = Tile<BY>::kSY, kThreads = Tile<BY>::kThreads;
    __shared__ float sM[3][kSY][kSX + 1];
    __shared__ float hsum[3][kSY][kBX + 1];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * kBX, y0 = blockIdx.y * kBY;
    const size_t plane = (size_t)H * W, chw = (size_t)C * plane;
    const int tid = threadIdx.x;
#pragma unroll
    for (int a = 0; a < 3; a++) stage_tile<kSY, kThreads>(maps + (size_t)a * chw + (size_t)c * plane, x0, y0, W, H, sM[a]);
    stage_wait();
    if (tid < kSY * (kBX / kHO)) {
        const int row = tid / (kBX / kHO), c0 = (tid % (kBX / kHO)) * kHO;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float p[kHO + kWin - 1];
#pragma unroll
            for (int j = 0; j < kHO + kWin - 1; j++) p[j] = sM[a][row][c0 + j];
#pragma unroll
            for (int o = 0; o < kHO; o++) {
                float acc = 0;
#pragma unroll
                for (int k = 0; k < kWin; k++) acc = fmaf(c_win[k], p[o + k], acc);
                hsum[a][row][c0 + o] = acc;
            }
        }
    }
    __syncthreads();
    const int tx = tid % kBX, r0 = (tid / kBX) * kVO;
    float v1[kVO], v2[kVO], v3[kVO];
#pragma unroll
    for (int o = 0; o < kVO; o++) { v1[o] = 0; v2[o] = 0; v3[o] = 0; }
#pragma unroll
    for (int j = 0; j < kVO + kWin - 1; j++) {
        const float h0 = hsum[0][r0 + j][tx], h1 = hsum[1][r0 + j][tx], h2 = hsum[2][r0 + j][tx];
#pragma unroll
        for (int o = 0; o < kVO; o++) {
            const int k = j - o;
            if (k >= 0 && k < kWin) {
                const float w = c_win[k];
                v1[o] = fmaf(w, h0, v1[o]); v2[o] = fmaf(w, h1, v2[o]); v3[o] = fmaf(w, h2, v3[o]);
            }
        }
    }
    const int gx = x0 + tx;
    const float g = grad_scale ? __ldg(grad_scale) : 1.0f;
#pragma unroll
    for (int o = 0; o < kVO; o++) {
        const int gy = y0 + r0 + o;
        if (gx < W && gy < H) {
            const size_t off = (size_t)c * plane + (size_t)gy * W + gx;
            const float p = __ldg(img + off), q = __ldg(gt + off);
            const float diff = p - q;
            const float sgn = diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f);  // torch.abs' subgradient: sign(0) = 0
            dimg[off] = g * inv_n * ((1.0f - lambda) * sgn - lambda * (v1[o] + 2.0f * p * v2[o] + q * v3[o]));
        }
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
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, radii, grad2d, max_radii2D, grad_accum, denom); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

// Tile height: 32 rows (256-thread CTAs, less halo traffic; measured at 1080p: fwd 92 us, bwd 68 us) or 16 rows
// (128-thread CTAs, ~26 KB of shared memory; fwd 92 us, bwd 76 us).  ISR_PHOTO_TILE overrides (16 / 32).
static int photo_tile_rows() {
    static const int v = [] { const char* e = getenv("ISR_PHOTO_TILE"); const int t = e ? atoi(e) : 0; return t == 16 ? 16 : 32; }();
    return v;
}

size_t photometric_ws_bytes(int C, int H, int W) {
    const size_t chw = (size_t)C * H * W;
    const size_t blocks = (size_t)((W + kBX - 1) / kBX) * ((H + 15) / 16) * C;  // enough for either tile height
    return align_up(3 * chw * 4, 256) + align_up(blocks * 8, 256);
}

int launch_photometric_fwd(int C, int H, int W, const float* img, const float* gt, float lambda, void* ws, float* out,
                           cudaStream_t stream) {
    const size_t chw = (size_t)C * H * W;
    float* maps = static_cast<float*>(ws);
    float* partials = reinterpret_cast<float*>(static_cast<char*>(ws) + align_up(3 * chw * 4, 256));
    const int by = photo_tile_rows();
    const dim3 grid((W + kBX - 1) / kBX, (H + by - 1) / by, C);
    if (by == 32) { photometric_fwd_kernel<32><<<grid, Tile<32>::kThreads, 0, stream>>>(C, H, W, img, gt, maps, partials); note_launch(); }
    else { photometric_fwd_kernel<16><<<grid, Tile<16>::kThreads, 0, stream>>>(C, H, W, img, gt, maps, partials); note_launch(); }
    photometric_reduce_kernel<<<1, 1024, 0, stream>>>(partials, (int)(grid.x * grid.y * grid.z), 1.0f / (float)chw, lambda, out); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_photometric_bwd(int C, int H, int W, const float* img, const float* gt, float lambda, const void* ws,
                           const float* grad_scale, float* dimg, cudaStream_t stream) {
    const size_t chw = (size_t)C * H * W;
    const int by = photo_tile_rows();
    const dim3 grid((W + kBX - 1) / kBX, (H + by - 1) / by, C);
    const float* maps = static_cast<const float*>(ws);
    if (by == 32) { photometric_bwd_kernel<32><<<grid, Tile<32>::kThreads, 0, stream>>>(C, H, W, img, gt, maps, grad_scale, lambda, 1.0f / (float)chw, dimg); note_launch(); }
    else { photometric_bwd_kernel<16><<<grid, Tile<16>::kThreads, 0, stream>>>(C, H, W, img, gt, maps, grad_scale, lambda, 1.0f / (float)chw, dimg); note_launch(); }
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

}  // namespace isr
