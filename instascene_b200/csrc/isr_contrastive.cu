// isr_contrastive.cu -- sampled-pixel gather + fused ProtoNCE contrastive loss (forward and backward).
// Reference: utils/contrastive_utils.py:18-73 (about 35 torch kernels, 3 torch.unique sorts and >= 4 host
// syncs per call) and the boolean-mask gather of train_semantic.py:124-129.
//
// Labels arrive already shifted (valid ids 0..K-1, invalid < 0), so the reference's unique()/remap steps
// (contrastive_utils.py:43-50) reduce to "clusters with a zero count do not exist".
//   f^_i  = f_i / (|f_i| + 1e-9)            norm detached                       (:41)
//   u_k   = predef_u[k]  or  mean_{i in k} f^_i   (gradient flows through mean)  (:44-45, :54-58)
//   phi_k = clip(10 * sum_{i in k} |f^_i - u_k| / (n_k log(n_k + lambda)), .5, 1)  detached   (:60-66)
//   loss  = - sum_i log( exp(f^_i.u_{y_i}/phi_{y_i}) / (sum_k exp(f^_i.u_k/phi_k) + 1e-9) )       (:68-71)
// The [N,F]x[F,K] logits contraction is ~70 MFLOP at N=32768,K=64,F=16: it is evaluated in exact fp32 FFMA
// (u_k staged in shared memory), one sample per thread.
#include "isr_common.cuh"

namespace isr {

struct ContrastWs {
    size_t fhat, inv_norm, counts, usum, u, phisum, phi, coef, dU, total;
    ContrastWs(int N, int F, int K) {
        size_t o = 0;
        fhat = o;     o = align_up(o + (size_t)N * F * 4, 256);
        inv_norm = o; o = align_up(o + (size_t)N * 4, 256);
        counts = o;   o = align_up(o + (size_t)K * 4, 256);
        usum = o;     o = align_up(o + (size_t)K * F * 4, 256);
        phisum = o;   o = align_up(o + (size_t)K * 4, 256);   // counts..phisum are zero-filled together
        u = o;        o = align_up(o + (size_t)K * F * 4, 256);
        phi = o;      o = align_up(o + (size_t)K * 4, 256);
        coef = o;     o = align_up(o + (size_t)N * K * 4, 256);
        dU = o;       o = align_up(o + (size_t)K * F * 4, 256);
        total = o;
    }
};

__global__ void gather_pixels_kernel(int F, int64_t HW, const float* __restrict__ map, int n,
                                     const int* __restrict__ pix_ids, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * F) return;
    const int s = i / F, ch = i - s * F;
    const int pix = pix_ids[s];
    out[i] = (pix >= 0 && pix < HW) ? map[(size_t)ch * HW + pix] : 0.0f;
}

// pass 1: normalise, count, (optionally) accumulate cluster sums
__global__ void contrast_normalise_kernel(int N, int F, int K, const float* __restrict__ feat,
                                          const int* __restrict__ labels, bool need_means, float* __restrict__ fhat,
                                          float* __restrict__ inv_norm, int* __restrict__ counts,
                                          float* __restrict__ usum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int y = labels[i];
    float ss = 0.0f;
    for (int c = 0; c < F; c++) { const float v = feat[(size_t)i * F + c]; ss = fmaf(v, v, ss); }
    const float inv = 1.0f / (sqrtf(ss) + 1e-9f);
    inv_norm[i] = inv;
    for (int c = 0; c < F; c++) fhat[(size_t)i * F + c] = feat[(size_t)i * F + c] * inv;
    if (y >= 0 && y < K) {
        atomicAdd(counts + y, 1);
        if (need_means)
            for (int c = 0; c < F; c++) atomicAdd(usum + (size_t)y * F + c, feat[(size_t)i * F + c] * inv);
    }
}

__global__ void contrast_means_kernel(int F, int K, const int* __restrict__ counts, const float* __restrict__ usum,
                                      const float* __restrict__ predef_u, float* __restrict__ u) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * F) return;
    const int k = i / F;
    if (predef_u) u[i] = predef_u[i];
    else u[i] = counts[k] > 0 ? usum[i] / (float)counts[k] : 0.0f;
}

__global__ void contrast_phisum_kernel(int N, int F, int K, const float* __restrict__ fhat,
                                       const int* __restrict__ labels, const float* __restrict__ u,
                                       float* __restrict__ phisum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int y = labels[i];
    if (y < 0 || y >= K) return;
    float ss = 0.0f;
    for (int c = 0; c < F; c++) { const float d = fhat[(size_t)i * F + c] - u[(size_t)y * F + c]; ss = fmaf(d, d, ss); }
    atomicAdd(phisum + y, sqrtf(ss));
}

__global__ void contrast_phi_kernel(int K, float temp_lambda, const int* __restrict__ counts,
                                    const float* __restrict__ phisum, float* __restrict__ phi) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const float n = (float)counts[k];
    float p = 1.0f;
    if (counts[k] > 0) p = fminf(fmaxf(10.0f * (phisum[k] / (n * logf(n + temp_lambda))), 0.5f), 1.0f);
    phi[k] = p;
}

// loss + softmax coefficients  coef[i][k] = (p_ik - [k == y_i]) / phi_k   (0 for absent clusters / invalid samples)
__global__ void __launch_bounds__(256)
contrast_loss_kernel(int N, int F, int K, const float* __restrict__ fhat, const int* __restrict__ labels,
                     const float* __restrict__ u, const float* __restrict__ phi, const int* __restrict__ counts,
                     float* __restrict__ coef, float* __restrict__ loss) {
    extern __shared__ float s_u[];  // [K][F] then phi[K] (0 marks an absent cluster)
    float* s_iphi = s_u + (size_t)K * F;
    for (int i = threadIdx.x; i < K * F; i += blockDim.x) s_u[i] = u[i];
    for (int k = threadIdx.x; k < K; k += blockDim.x) s_iphi[k] = counts[k] > 0 ? phi[k] : 0.0f;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float li = 0.0f;
    if (i < N) {
        const int y = labels[i];
        const bool valid = (y >= 0 && y < K);
        float f[ISR_MAX_EXTRA_DIMS];
        for (int c = 0; c < F; c++) f[c] = fhat[(size_t)i * F + c];
        float sum = 0.0f, dy = 0.0f;
        if (valid) {
            for (int k = 0; k < K; k++) {
                if (s_iphi[k] == 0.0f) continue;
                float dot = 0.0f;
                for (int c = 0; c < F; c++) dot = fmaf(f[c], s_u[(size_t)k * F + c], dot);
                const float e = expf(dot / s_iphi[k]);
                sum += e;
                if (k == y) dy = e;
                coef[(size_t)i * K + k] = e;
            }
            const float denom = sum + 1e-9f;
            li = -logf(dy / denom);
            for (int k = 0; k < K; k++) {
                const float ip = s_iphi[k];
                const float p = ip == 0.0f ? 0.0f : coef[(size_t)i * K + k] / denom;
                coef[(size_t)i * K + k] = ip == 0.0f ? 0.0f : (p - (k == y ? 1.0f : 0.0f)) / ip;
            }
        } else {
            for (int k = 0; k < K; k++) coef[(size_t)i * K + k] = 0.0f;
        }
    }
    // block reduction of the loss
    __shared__ float s_red[8];
    for (int off = 16; off > 0; off >>= 1) li += __shfl_xor_sync(0xffffffffu, li, off);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = li;
    __syncthreads();
    if (threadIdx.x < 8) {
        float v = s_red[threadIdx.x];
        for (int off = 4; off > 0; off >>= 1) v += __shfl_xor_sync(0xffu, v, off);
        if (threadIdx.x == 0) atomicAdd(loss, v);
    }
}

// dU[k][c] = sum_i coef[i][k] * fhat[i][c]: each block reduces a slab of 64 samples for all K*F entries (entry e =
// k*F + c owned by thread e mod 256), then adds its partial sums to dU with one atomic per entry.
__global__ void __launch_bounds__(256)
contrast_dU_kernel(int N, int F, int K, const float* __restrict__ coef, const float* __restrict__ fhat,
                   float* __restrict__ dU) {
    const int i0 = blockIdx.x * 64, i1 = min(N, i0 + 64);
    const int KF = K * F;
    for (int e = threadIdx.x; e < KF; e += 256) {
        const int k = e / F, cch = e - k * F;
        float acc = 0.0f;
        for (int i = i0; i < i1; i++) acc = fmaf(__ldg(coef + (size_t)i * K + k), __ldg(fhat + (size_t)i * F + cch), acc);
        if (acc != 0.0f) atomicAdd(dU + e, acc);
    }
}

__global__ void __launch_bounds__(256)
contrast_dfeat_kernel(int N, int F, int K, const float* __restrict__ coef, const int* __restrict__ labels,
                      const float* __restrict__ u, const float* __restrict__ dU, const int* __restrict__ counts,
                      const float* __restrict__ inv_norm, bool means, const float* __restrict__ grad_scale,
                      float* __restrict__ dfeat) {
    extern __shared__ float s_u[];  // [K][F]
    for (int i = threadIdx.x; i < K * F; i += blockDim.x) s_u[i] = u[i];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int y = labels[i];
    float g[ISR_MAX_EXTRA_DIMS];
    for (int c = 0; c < F; c++) g[c] = 0.0f;
    if (y >= 0 && y < K) {
        for (int k = 0; k < K; k++) {
            const float cf = coef[(size_t)i * K + k];
            if (cf != 0.0f)
                for (int c = 0; c < F; c++) g[c] = fmaf(cf, s_u[(size_t)k * F + c], g[c]);
        }
        if (means) {
            const float inv_n = 1.0f / (float)counts[y];
            for (int c = 0; c < F; c++) g[c] = fmaf(dU[(size_t)y * F + c], inv_n, g[c]);
        }
    }
    const float sc = (grad_scale ? *grad_scale : 1.0f) * inv_norm[i];
    for (int c = 0; c < F; c++) dfeat[(size_t)i * F + c] = g[c] * sc;
}

// ---- fused row normalisation (one thread per row, F <= 32 values in registers) ---------------------------------
template <int FP>
__global__ void __launch_bounds__(256)
rownorm_fwd_kernel(int P, int F, const float* __restrict__ x, float eps1, float eps2, int stages, float* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float v[FP];
    const bool vec = (F & 3) == 0;
#pragma unroll
    for (int c = 0; c < FP; c += 4) {
        if (vec && c < F) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(x + (size_t)i * F + c));
            v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) v[c + k] = (c + k < F) ? __ldg(x + (size_t)i * F + c + k) : 0.0f;
        }
    }
    float ss = 0.0f;
#pragma unroll
    for (int c = 0; c < FP; c++) ss = fmaf(v[c], v[c], ss);
    float inv = 1.0f / (sqrtf(ss) + eps1);
#pragma unroll
    for (int c = 0; c < FP; c++) v[c] *= inv;
    if (stages > 1) {
        ss = 0.0f;
#pragma unroll
        for (int c = 0; c < FP; c++) ss = fmaf(v[c], v[c], ss);
        inv = 1.0f / (sqrtf(ss) + eps2);
#pragma unroll
        for (int c = 0; c < FP; c++) v[c] *= inv;
    }
#pragma unroll
    for (int c = 0; c < FP; c += 4) {
        if (vec && c < F) {
            *reinterpret_cast<float4*>(y + (size_t)i * F + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) if (c + k < F) y[(size_t)i * F + c + k] = v[c + k];
        }
    }
}

// d/dx of y = x / (n + eps), n = |x|:  dx = dy/(n+eps) - x (x.dy) / (n (n+eps)^2)
template <int FP>
__device__ __forceinline__ void rownorm_vjp(const float* x, float* d, float eps) {
    float ss = 0.0f, xd = 0.0f;
#pragma unroll
    for (int c = 0; c < FP; c++) { ss = fmaf(x[c], x[c], ss); xd = fmaf(x[c], d[c], xd); }
    const float n = sqrtf(ss), ne = n + eps;
    const float a = 1.0f / ne;
    const float b = n > 0.0f ? xd / (n * ne * ne) : 0.0f;
#pragma unroll
    for (int c = 0; c < FP; c++) d[c] = fmaf(-b, x[c], a * d[c]);
}

template <int FP>
__global__ void __launch_bounds__(256)
rownorm_bwd_kernel(int P, int F, const float* __restrict__ x, const float* __restrict__ dy, float eps1, float eps2,
                   int stages, float* __restrict__ dx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float v[FP], d[FP];
    const bool vec = (F & 3) == 0;
#pragma unroll
    for (int c = 0; c < FP; c += 4) {
        if (vec && c < F) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(x + (size_t)i * F + c));
            const float4 u = __ldg(reinterpret_cast<const float4*>(dy + (size_t)i * F + c));
            v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
            d[c] = u.x; d[c + 1] = u.y; d[c + 2] = u.z; d[c + 3] = u.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                v[c + k] = (c + k < F) ? __ldg(x + (size_t)i * F + c + k) : 0.0f;
                d[c + k] = (c + k < F) ? __ldg(dy + (size_t)i * F + c + k) : 0.0f;
            }
        }
    }
    if (stages > 1) {
        float a1[FP];
        float ss = 0.0f;
#pragma unroll
        for (int c = 0; c < FP; c++) ss = fmaf(v[c], v[c], ss);
        const float inv = 1.0f / (sqrtf(ss) + eps1);
#pragma unroll
        for (int c = 0; c < FP; c++) a1[c] = v[c] * inv;
        rownorm_vjp<FP>(a1, d, eps2);
    }
    rownorm_vjp<FP>(v, d, eps1);
#pragma unroll
    for (int c = 0; c < FP; c += 4) {
        if (vec && c < F) {
            *reinterpret_cast<float4*>(dx + (size_t)i * F + c) = make_float4(d[c], d[c + 1], d[c + 2], d[c + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) if (c + k < F) dx[(size_t)i * F + c + k] = d[c + k];
        }
    }
}

template <int FP>
static int rownorm_launch(bool fwd, int P, int F, const float* x, const float* dy, float e1, float e2, int stages,
                          float* out, cudaStream_t stream) {
    if (fwd) rownorm_fwd_kernel<FP><<<(P + 255) / 256, 256, 0, stream>>>(P, F, x, e1, e2, stages, out);
    else rownorm_bwd_kernel<FP><<<(P + 255) / 256, 256, 0, stream>>>(P, F, x, dy, e1, e2, stages, out);
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_rownorm(bool fwd, int P, int F, const float* x, const float* dy, float e1, float e2, int stages, float* out,
                   cudaStream_t stream) {
    if (P <= 0 || F <= 0) return ISR_OK;
    if (F <= 4) return rownorm_launch<4>(fwd, P, F, x, dy, e1, e2, stages, out, stream);
    if (F <= 8) return rownorm_launch<8>(fwd, P, F, x, dy, e1, e2, stages, out, stream);
    if (F <= 16) return rownorm_launch<16>(fwd, P, F, x, dy, e1, e2, stages, out, stream);
    if (F <= 24) return rownorm_launch<24>(fwd, P, F, x, dy, e1, e2, stages, out, stream);
    if (F <= 32) return rownorm_launch<32>(fwd, P, F, x, dy, e1, e2, stages, out, stream);
    return ISR_ERR_UNSUPPORTED;
}

size_t contrastive_ws_bytes(int N, int F, int K) { return ContrastWs(N, F, K).total; }

int launch_gather_pixels(int F, int64_t HW, const float* map, int n, const int* pix_ids, float* out, cudaStream_t stream) {
    if (n <= 0 || F <= 0) return ISR_OK;
    const int total = n * F;
    gather_pixels_kernel<<<(total + 255) / 256, 256, 0, stream>>>(F, HW, map, n, pix_ids, out);
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_contrastive_fwd(int N, int F, int K, const float* features, const int* labels, const float* predef_u,
                           float temp_lambda, void* ws, float* loss, cudaStream_t stream) {
    ContrastWs L(N, F, K);
    char* w = static_cast<char*>(ws);
    ISR_CUDA_TRY(cudaMemsetAsync(w + L.counts, 0, L.u - L.counts, stream));
    ISR_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), stream));
    if (N <= 0 || K <= 0) return ISR_OK;
    float* fhat = reinterpret_cast<float*>(w + L.fhat);
    int* counts = reinterpret_cast<int*>(w + L.counts);
    float* u = reinterpret_cast<float*>(w + L.u);
    float* phi = reinterpret_cast<float*>(w + L.phi);
    contrast_normalise_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, F, K, features, labels, predef_u == nullptr, fhat,
                                                                   reinterpret_cast<float*>(w + L.inv_norm), counts,
                                                                   reinterpret_cast<float*>(w + L.usum));
    contrast_means_kernel<<<(K * F + 255) / 256, 256, 0, stream>>>(F, K, counts, reinterpret_cast<float*>(w + L.usum),
                                                                   predef_u, u);
    contrast_phisum_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, F, K, fhat, labels, u,
                                                                reinterpret_cast<float*>(w + L.phisum));
    contrast_phi_kernel<<<(K + 255) / 256, 256, 0, stream>>>(K, temp_lambda, counts,
                                                             reinterpret_cast<float*>(w + L.phisum), phi);
    const size_t smem = ((size_t)K * F + K) * sizeof(float);
    ISR_CUDA_TRY(cudaFuncSetAttribute(contrast_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    contrast_loss_kernel<<<(N + 255) / 256, 256, smem, stream>>>(N, F, K, fhat, labels, u, phi, counts,
                                                                 reinterpret_cast<float*>(w + L.coef), loss);
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_contrastive_bwd(int N, int F, int K, const int* labels, const float* predef_u, const void* ws,
                           const float* grad_scale, float* dfeat, cudaStream_t stream) {
    if (N <= 0) return ISR_OK;
    ContrastWs L(N, F, K);
    const char* w = static_cast<const char*>(ws);
    const bool means = predef_u == nullptr;
    float* dU = reinterpret_cast<float*>(const_cast<char*>(w) + L.dU);
    if (means && K > 0) {
        ISR_CUDA_TRY(cudaMemsetAsync(dU, 0, (size_t)K * F * sizeof(float), stream));
        contrast_dU_kernel<<<(N + 63) / 64, 256, 0, stream>>>(N, F, K, reinterpret_cast<const float*>(w + L.coef),
                                                                reinterpret_cast<const float*>(w + L.fhat), dU);
    }
    const size_t smem = (size_t)(K > 0 ? K : 1) * F * sizeof(float);
    ISR_CUDA_TRY(cudaFuncSetAttribute(contrast_dfeat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    contrast_dfeat_kernel<<<(N + 255) / 256, 256, smem, stream>>>(
        N, F, K, reinterpret_cast<const float*>(w + L.coef), labels, reinterpret_cast<const float*>(w + L.u), dU,
        reinterpret_cast<const int*>(w + L.counts), reinterpret_cast<const float*>(w + L.inv_norm), means, grad_scale,
        dfeat);
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

}  // namespace isr
