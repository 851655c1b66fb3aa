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
#include <cmath>

#include "isr_common.cuh"

namespace isr {

// Workspace: f^ [N,F], 1/(|f|+eps) [N], counts [K], u [K,F], phi [K], coefT [K,N] (softmax coefficients, CLUSTER-major
// so that both the per-cluster reductions and the per-sample reads are coalesced), dU [K,F].
struct ContrastWs {
    size_t fhat, inv_norm, counts, u, phi, coef, dU, total;
    ContrastWs(int N, int F, int K) {
        size_t o = 0;
        fhat = o;     o = align_up(o + (size_t)N * F * 4, 256);
        inv_norm = o; o = align_up(o + (size_t)N * 4, 256);
        counts = o;   o = align_up(o + (size_t)K * 4, 256);
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

// 1. per sample: f^ = f / (|f| + 1e-9)
__global__ void contrast_normalise_kernel(int N, int F, const float* __restrict__ feat, float* __restrict__ fhat,
                                          float* __restrict__ inv_norm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float ss = 0.0f;
    for (int c = 0; c < F; c++) { const float v = feat[(size_t)i * F + c]; ss = fmaf(v, v, ss); }
    const float inv = 1.0f / (sqrtf(ss) + 1e-9f);
    inv_norm[i] = inv;
    for (int c = 0; c < F; c++) fhat[(size_t)i * F + c] = feat[(size_t)i * F + c] * inv;
}

// sum over the whole block (up to 1024 threads); s_red must hold 32 floats
__device__ __forceinline__ float block_sum(float v, float* s_red) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; w++) t += s_red[w];
    return t;
}

// 2. one block per cluster k: count, centre u_k (mean of members or predefined prototype), temperature phi_k.
//    No atomics, deterministic; the label vector (N ints) is re-read by every block from L2.
__global__ void __launch_bounds__(1024)
contrast_cluster_kernel(int N, int F, int K, const float* __restrict__ fhat, const int* __restrict__ labels,
                        const float* __restrict__ predef_u, float temp_lambda, int* __restrict__ counts,
                        float* __restrict__ u, float* __restrict__ phi) {
    __shared__ float s_red[32];
    __shared__ float s_u[ISR_MAX_EXTRA_DIMS];
    const int k = blockIdx.x;
    float cnt = 0.0f, acc[ISR_MAX_EXTRA_DIMS];
    for (int c = 0; c < F; c++) acc[c] = 0.0f;
    for (int i = threadIdx.x; i < N; i += blockDim.x)
        if (__ldg(labels + i) == k) {
            cnt += 1.0f;
            if (predef_u == nullptr)
                for (int c = 0; c < F; c++) acc[c] += fhat[(size_t)i * F + c];
        }
    const float n = block_sum(cnt, s_red);
    for (int c = 0; c < F; c++) {
        float m;
        if (predef_u != nullptr) m = predef_u[(size_t)k * F + c];
        else { const float t = block_sum(acc[c], s_red); m = n > 0.0f ? t / n : 0.0f; }
        if (threadIdx.x == 0) { s_u[c] = m; u[(size_t)k * F + c] = m; }
    }
    __syncthreads();
    float spread = 0.0f;
    for (int i = threadIdx.x; i < N; i += blockDim.x)
        if (__ldg(labels + i) == k) {
            float ss = 0.0f;
            for (int c = 0; c < F; c++) { const float d = fhat[(size_t)i * F + c] - s_u[c]; ss = fmaf(d, d, ss); }
            spread += sqrtf(ss);
        }
    const float tot = block_sum(spread, s_red);
    if (threadIdx.x == 0) {
        counts[k] = (int)n;
        phi[k] = n > 0.0f ? fminf(fmaxf(10.0f * (tot / (n * logf(n + temp_lambda))), 0.5f), 1.0f) : 1.0f;
    }
}

// 3. loss + softmax coefficients coefT[k][i] = (p_ik - [k == y_i]) / phi_k (0 for absent clusters / ignored samples).
//    Four lanes per sample, each covering a quarter of the clusters; u and phi staged in shared memory.
__global__ void __launch_bounds__(256)
contrast_loss_kernel(int N, int F, int K, const float* __restrict__ fhat, const int* __restrict__ labels,
                     const float* __restrict__ u, const float* __restrict__ phi, const int* __restrict__ counts,
                     float* __restrict__ coefT, float* __restrict__ loss) {
    extern __shared__ float s_u[];  // [K][F] then phi[K] (0 marks an absent cluster)
    float* s_phi = s_u + (size_t)K * F;
    for (int i = threadIdx.x; i < K * F; i += blockDim.x) s_u[i] = u[i];
    for (int k = threadIdx.x; k < K; k += blockDim.x) s_phi[k] = counts[k] > 0 ? phi[k] : 0.0f;
    __syncthreads();
    const int q = threadIdx.x & 3;                          // quarter of the clusters
    const int i = blockIdx.x * 64 + (threadIdx.x >> 2);     // sample
    const int kq = (K + 3) / 4, k0 = q * kq, k1 = min(K, k0 + kq);
    float li = 0.0f;
    const bool in_range = i < N;
    const int y = in_range ? labels[i] : -1;
    const bool valid = in_range && y >= 0 && y < K;
    float f[ISR_MAX_EXTRA_DIMS];
    if (in_range) for (int c = 0; c < F; c++) f[c] = fhat[(size_t)i * F + c];
    float sum = 0.0f, dy = 0.0f;
    if (valid)
        for (int k = k0; k < k1; k++) {
            const float ph = s_phi[k];
            float e = 0.0f;
            if (ph != 0.0f) {
                float dot = 0.0f;
                for (int c = 0; c < F; c++) dot = fmaf(f[c], s_u[(size_t)k * F + c], dot);
                e = expf(dot / ph);
                sum += e;
                if (k == y) dy = e;
            }
            coefT[(size_t)k * N + i] = e;
        }
    // combine the four quarters (lanes 4j..4j+3 of a warp belong to one sample)
    sum += __shfl_xor_sync(0xffffffffu, sum, 1); sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    dy += __shfl_xor_sync(0xffffffffu, dy, 1);   dy += __shfl_xor_sync(0xffffffffu, dy, 2);
    if (in_range) {
        const float denom = sum + 1e-9f;
        if (valid && q == 0) li = -logf(dy / denom);
        for (int k = k0; k < k1; k++) {
            const float ph = s_phi[k];
            float cf = 0.0f;
            if (valid && ph != 0.0f) cf = (coefT[(size_t)k * N + i] / denom - (k == y ? 1.0f : 0.0f)) / ph;
            coefT[(size_t)k * N + i] = cf;
        }
    }
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

// 4. dU[k][c] = sum_i coefT[k][i] * f^[i][c]: split over slabs of 128 samples; a thread owns the entries e = k*F + c with
//    e = tid (mod 256): its coefficient reads run along the contiguous sample axis of coefT (16-byte loads, shared by
//    the F threads of a cluster), its feature reads are coalesced across c.  One atomic per entry per slab.
__global__ void __launch_bounds__(256)
contrast_dU_kernel(int N, int F, int K, const float* __restrict__ coefT, const float* __restrict__ fhat,
                   float* __restrict__ dU) {
    constexpr int SLAB = 128;
    const int i0 = blockIdx.x * SLAB, i1 = min(N, i0 + SLAB);
    const int KF = K * F;
    const bool vec = ((N & 3) == 0) && (i1 - i0 == SLAB);
    for (int e = threadIdx.x; e < KF; e += 256) {
        const int k = e / F, cch = e - k * F;
        const float* cr = coefT + (size_t)k * N;
        float acc = 0.0f;
        if (vec) {
            for (int i = i0; i < i1; i += 4) {
                const float4 cf = __ldg(reinterpret_cast<const float4*>(cr + i));
                acc = fmaf(cf.x, __ldg(fhat + (size_t)i * F + cch), acc);
                acc = fmaf(cf.y, __ldg(fhat + (size_t)(i + 1) * F + cch), acc);
                acc = fmaf(cf.z, __ldg(fhat + (size_t)(i + 2) * F + cch), acc);
                acc = fmaf(cf.w, __ldg(fhat + (size_t)(i + 3) * F + cch), acc);
            }
        } else {
            for (int i = i0; i < i1; i++) acc = fmaf(__ldg(cr + i), __ldg(fhat + (size_t)i * F + cch), acc);
        }
        if (acc != 0.0f) atomicAdd(dU + e, acc);
    }
}

// 5. dL/df_i = grad_scale / (|f_i| + eps) * ( sum_k coefT[k][i] u_k  +  [means] dU[y_i] / n_{y_i} )
__global__ void __launch_bounds__(256)
contrast_dfeat_kernel(int N, int F, int K, const float* __restrict__ coefT, const int* __restrict__ labels,
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
            const float cf = __ldg(coefT + (size_t)k * N + i);
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

// ---- fused Adam step (one pass over param / grad / exp_avg / exp_avg_sq; torch.optim.Adam semantics, no amsgrad) ----
__global__ void __launch_bounds__(256)
adam_step_kernel(size_t n4, size_t n, float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                 float4* __restrict__ v, float beta1, float beta2, float eps, float step_size, float inv_sqrt_bias2) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        mm = beta1 * mm + (1.0f - beta1) * gg;
        vv = beta2 * vv + (1.0f - beta2) * gg * gg;
        pp -= step_size * mm / (sqrtf(vv) * inv_sqrt_bias2 + eps);
    };
    if (i < n4) {
        float4 P = p[i], M = m[i], V = v[i];
        const float4 G = g[i];
        upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
        p[i] = P; m[i] = M; v[i] = V;
    } else if (i == n4) {  // tail (n not a multiple of 4)
        float* ps = reinterpret_cast<float*>(p); const float* gs = reinterpret_cast<const float*>(g);
        float* ms = reinterpret_cast<float*>(m); float* vs = reinterpret_cast<float*>(v);
        for (size_t j = n4 * 4; j < n; j++) upd(ps[j], gs[j], ms[j], vs[j]);
    }
}

int launch_adam(size_t n, float* p, const float* g, float* m, float* v, float lr, float beta1, float beta2, float eps,
                int step, cudaStream_t stream) {
    if (n == 0) return ISR_OK;
    const double b1 = 1.0 - pow((double)beta1, (double)step), b2 = 1.0 - pow((double)beta2, (double)step);
    const size_t n4 = n / 4;
    adam_step_kernel<<<(unsigned)((n4 + 1 + 255) / 256), 256, 0, stream>>>(
        n4, n, reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
        reinterpret_cast<float4*>(v), beta1, beta2, eps, (float)((double)lr / b1), (float)(1.0 / sqrt(b2)));
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
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
    ISR_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), stream));
    if (N <= 0 || K <= 0) return ISR_OK;
    float* fhat = reinterpret_cast<float*>(w + L.fhat);
    int* counts = reinterpret_cast<int*>(w + L.counts);
    float* u = reinterpret_cast<float*>(w + L.u);
    float* phi = reinterpret_cast<float*>(w + L.phi);
    contrast_normalise_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, F, features, fhat, reinterpret_cast<float*>(w + L.inv_norm));
    contrast_cluster_kernel<<<K, 1024, 0, stream>>>(N, F, K, fhat, labels, predef_u, temp_lambda, counts, u, phi);
    const size_t smem = ((size_t)K * F + K) * sizeof(float);
    ISR_CUDA_TRY(cudaFuncSetAttribute(contrast_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    contrast_loss_kernel<<<(N + 63) / 64, 256, smem, stream>>>(N, F, K, fhat, labels, u, phi, counts,
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
        contrast_dU_kernel<<<(N + 127) / 128, 256, 0, stream>>>(N, F, K, reinterpret_cast<const float*>(w + L.coef),
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
