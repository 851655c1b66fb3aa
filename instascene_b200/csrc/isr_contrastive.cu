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

// Workspace.  `zeroed` (cluster sums [K,F], counts [K], spreads [K], dU [K,F]) is cleared once per forward.
struct ContrastWs {
    size_t fhat, inv_norm, g, zeroed, sums, counts, spread, dU, zeroed_bytes, total;
    ContrastWs(int N, int F, int K) {
        size_t o = 0;
        fhat = o;     o = align_up(o + (size_t)N * F * 4, 256);
        inv_norm = o; o = align_up(o + (size_t)N * 4, 256);
        g = o;        o = align_up(o + (size_t)N * F * 4, 256);   // sum_k coef_ik u_k (backward, without the dU term)
        zeroed = o;
        sums = o;     o = align_up(o + (size_t)K * F * 4, 16);
        counts = o;   o = align_up(o + (size_t)K * 4, 16);
        spread = o;   o = align_up(o + (size_t)K * 4, 16);
        dU = o;       o = align_up(o + (size_t)K * F * 4, 256);
        zeroed_bytes = o - zeroed;
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

// The loss is evaluated in four grid-wide phases (each needs a reduction over ALL samples of the previous one), every
// phase one kernel of ceil(N/256) blocks x 256 threads, one sample per thread:
//   stats   f^ = f/(|f|+1e-9); per-cluster sums and counts
//   spread  u_k = mean (or predefined prototype); per-cluster sum of |f^ - u_k|
//   loss    phi_k; logits, loss, softmax coefficients coef_ik = (p_ik - [k == y_i]) / phi_k (never stored: a block keeps
//           64 clusters x 256 samples of them in shared memory), g_i = sum_k coef_ik u_k, dU_k = sum_i coef_ik f^_i
//   dfeat   (backward) dL/df_i = grad_scale / (|f_i| + eps) * (g_i + [means] dU[y_i] / n_{y_i})
// Per-cluster block reductions are "owner computes": the block's samples and labels sit in shared memory and a thread
// sums the members of the (cluster, channel) entries it owns -- no shared-memory float atomics (CAS loops on this
// target); one global red per entry per block.
constexpr int kCB = 256;   // samples per block
constexpr int kKC = 64;    // clusters per coefficient chunk

// sum over the block (kCB threads); s_red must hold kCB/32 floats
__device__ __forceinline__ float block_sum(float v, float* s_red) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    for (int w = 0; w < kCB / 32; w++) t += s_red[w];
    return t;
}

// block partial of  out[k][c] += sum_{i in block, label_i == k} w_i * sf[i][c]   for k in [k0, k0+nk)
// (w == nullptr: w_i = 1).  sw is indexed [k - k0][i] when per-cluster weights are given (kPerCluster).
template <bool kPerCluster>
__device__ __forceinline__ void owner_accumulate(int F, int FS, int k0, int nk, const float* __restrict__ sf,
                                                 const int* __restrict__ sl, const float* __restrict__ sw,
                                                 float* __restrict__ out) {
    if ((F & 3) == 0) {
        const int F4 = F >> 2;
        for (int e = threadIdx.x; e < nk * F4; e += kCB) {
            const int kk = e / F4, c4 = e - kk * F4, k = k0 + kk;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = 0; i < kCB; i++) {
                float w;
                if (kPerCluster) { w = sw[kk * kCB + i]; if (w == 0.0f) continue; }
                else { if (sl[i] != k) continue; w = 1.0f; }
                const float4 f = *reinterpret_cast<const float4*>(sf + i * FS + 4 * c4);
                acc.x = fmaf(w, f.x, acc.x); acc.y = fmaf(w, f.y, acc.y); acc.z = fmaf(w, f.z, acc.z); acc.w = fmaf(w, f.w, acc.w);
            }
            float* o = out + (size_t)k * F + 4 * c4;
            if (acc.x != 0.0f) atomicAdd(o + 0, acc.x);
            if (acc.y != 0.0f) atomicAdd(o + 1, acc.y);
            if (acc.z != 0.0f) atomicAdd(o + 2, acc.z);
            if (acc.w != 0.0f) atomicAdd(o + 3, acc.w);
        }
    } else {
        for (int e = threadIdx.x; e < nk * F; e += kCB) {
            const int kk = e / F, c = e - kk * F, k = k0 + kk;
            float acc = 0.0f;
            for (int i = 0; i < kCB; i++) {
                float w;
                if (kPerCluster) { w = sw[kk * kCB + i]; if (w == 0.0f) continue; }
                else { if (sl[i] != k) continue; w = 1.0f; }
                acc = fmaf(w, sf[i * FS + c], acc);
            }
            if (acc != 0.0f) atomicAdd(out + (size_t)k * F + c, acc);
        }
    }
}

// shared-memory row stride of the block's sample matrix (16-byte aligned rows, odd multiple of 4 words: the owner
// loops read one row at a time, so there is nothing to de-conflict beyond alignment)
__host__ __device__ inline int sample_stride(int F) { return (F + 3) & ~3; }

// phase 1   (FP: F padded to 4/8/16/24/32 so that the per-sample vectors stay in registers)
template <int FP>
__global__ void __launch_bounds__(kCB)
contrast_stats_kernel(int N, int F, int K, const float* __restrict__ feat, const int* __restrict__ labels, bool want_sums,
                      float* __restrict__ fhat, float* __restrict__ inv_norm, float* __restrict__ sums,
                      float* __restrict__ counts, float* __restrict__ loss) {
    extern __shared__ __align__(16) float smem[];
    const int FS = sample_stride(F);
    float* sf = smem;                                        // [kCB][FS]
    int* sl = reinterpret_cast<int*>(smem + kCB * FS);       // [kCB]
    const int i = blockIdx.x * kCB + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) *loss = 0.0f;   // accumulated by phase 3
    int y = -1;
    if (i < N) {
        y = labels[i];
        if (y < 0 || y >= K) y = -1;
        float v[FP], ss = 0.0f;
#pragma unroll
        for (int c = 0; c < FP; c++) { v[c] = c < F ? feat[(size_t)i * F + c] : 0.0f; ss = fmaf(v[c], v[c], ss); }
        const float inv = 1.0f / (sqrtf(ss) + 1e-9f);
        inv_norm[i] = inv;
#pragma unroll
        for (int c = 0; c < FP; c++)
            if (c < F) { const float h = v[c] * inv; fhat[(size_t)i * F + c] = h; sf[threadIdx.x * FS + c] = h; }
    }
    sl[threadIdx.x] = y;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += kCB) {
        int n = 0;
        for (int j = 0; j < kCB; j++) n += (sl[j] == k);
        if (n) atomicAdd(counts + k, (float)n);
    }
    if (want_sums) owner_accumulate<false>(F, FS, 0, K, sf, sl, nullptr, sums);
}

// u_k into shared memory: predefined prototype or mean of the members (0 for an absent cluster)
__device__ __forceinline__ void load_centres(int F, int K, const float* __restrict__ predef_u,
                                             const float* __restrict__ sums, const float* __restrict__ counts,
                                             float* __restrict__ su) {
    for (int e = threadIdx.x; e < K * F; e += kCB) {
        const int k = e / F;
        const float n = counts[k];
        su[e] = predef_u ? predef_u[e] : (n > 0.0f ? sums[e] / n : 0.0f);
    }
}

// phase 2
__global__ void __launch_bounds__(kCB)
contrast_spread_kernel(int N, int F, int K, const float* __restrict__ fhat, const int* __restrict__ labels,
                       const float* __restrict__ predef_u, const float* __restrict__ sums,
                       const float* __restrict__ counts, float* __restrict__ spread) {
    extern __shared__ __align__(16) float smem[];
    float* su = smem;                                   // [K][F]
    float* sd = smem + (size_t)K * F;                   // [kCB]
    int* sl = reinterpret_cast<int*>(sd + kCB);         // [kCB]
    load_centres(F, K, predef_u, sums, counts, su);
    __syncthreads();
    const int i = blockIdx.x * kCB + threadIdx.x;
    int y = -1;
    float d = 0.0f;
    if (i < N) {
        y = labels[i];
        if (y < 0 || y >= K) y = -1;
        if (y >= 0) {
            float ss = 0.0f;
            for (int c = 0; c < F; c++) { const float t = fhat[(size_t)i * F + c] - su[(size_t)y * F + c]; ss = fmaf(t, t, ss); }
            d = sqrtf(ss);
        }
    }
    sd[threadIdx.x] = d;
    sl[threadIdx.x] = y;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += kCB) {
        float s = 0.0f;
        for (int j = 0; j < kCB; j++) if (sl[j] == k) s += sd[j];
        if (s != 0.0f) atomicAdd(spread + k, s);
    }
}

// phase 3
template <int FP>
__global__ void __launch_bounds__(kCB)
contrast_loss_kernel(int N, int F, int K, const float* __restrict__ fhat, const int* __restrict__ labels,
                     const float* __restrict__ predef_u, const float* __restrict__ sums,
                     const float* __restrict__ counts, const float* __restrict__ spread, float temp_lambda,
                     float* __restrict__ g_out, float* __restrict__ dU, float* __restrict__ loss) {
    extern __shared__ __align__(16) float smem[];
    const int FS = sample_stride(F);
    float* su = smem;                                   // [K][F]
    float* sphi = su + (size_t)K * F;                   // [K]  1/phi_k, 0 for an absent cluster
    float* sf = sphi + ((K + 3) & ~3);                  // [kCB][FS]
    float* sw = sf + kCB * FS;                          // [kKC][kCB] coefficient chunk
    __shared__ float s_red[kCB / 32];
    const bool means = predef_u == nullptr;
    load_centres(F, K, predef_u, sums, counts, su);
    for (int k = threadIdx.x; k < K; k += kCB) {
        const float n = counts[k];
        sphi[k] = n > 0.0f ? 1.0f / fminf(fmaxf(10.0f * (spread[k] / (n * logf(n + temp_lambda))), 0.5f), 1.0f) : 0.0f;
    }
    const int i = blockIdx.x * kCB + threadIdx.x;
    int y = -1;
    float f[FP];
#pragma unroll
    for (int c = 0; c < FP; c++) f[c] = 0.0f;
    if (i < N) {
        y = labels[i];
        if (y < 0 || y >= K) y = -1;
#pragma unroll
        for (int c = 0; c < FP; c++)
            if (c < F) { f[c] = fhat[(size_t)i * F + c]; sf[threadIdx.x * FS + c] = f[c]; }
    }
    __syncthreads();
    const bool valid = y >= 0;
    auto logit_exp = [&](int k) {  // exp(f . u_k / phi_k), 0 for an absent cluster
        const float ip = sphi[k];
        if (ip == 0.0f) return 0.0f;
        float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
        for (int c = 0; c < FP; c += 2) {
            if (c < F) d0 = fmaf(f[c], su[(size_t)k * F + c], d0);
            if (c + 1 < F) d1 = fmaf(f[c + 1], su[(size_t)k * F + c + 1], d1);
        }
        return expf((d0 + d1) * ip);
    };
    float sum = 0.0f, dy = 0.0f;
    if (valid)
        for (int k = 0; k < K; k++) {
            const float e = logit_exp(k);
            sum += e;
            if (k == y) dy = e;
        }
    const float denom = sum + 1e-9f;
    const float li = block_sum(valid ? -logf(dy / denom) : 0.0f, s_red);
    if (threadIdx.x == 0 && li != 0.0f) atomicAdd(loss, li);
    const float inv_denom = 1.0f / denom;
    float g[FP];
#pragma unroll
    for (int c = 0; c < FP; c++) g[c] = 0.0f;
    for (int k0 = 0; k0 < K; k0 += kKC) {
        const int nk = min(kKC, K - k0);
        for (int kk = 0; kk < nk; kk++) {
            const int k = k0 + kk;
            float cf = 0.0f;
            if (valid && sphi[k] != 0.0f) {
                cf = (logit_exp(k) * inv_denom - (k == y ? 1.0f : 0.0f)) * sphi[k];
#pragma unroll
                for (int c = 0; c < FP; c++)
                    if (c < F) g[c] = fmaf(cf, su[(size_t)k * F + c], g[c]);
            }
            if (means) sw[kk * kCB + threadIdx.x] = cf;
        }
        if (means) {
            __syncthreads();
            owner_accumulate<true>(F, FS, k0, nk, sf, nullptr, sw, dU);
            __syncthreads();
        }
    }
    if (i < N) {
#pragma unroll
        for (int c = 0; c < FP; c++)
            if (c < F) g_out[(size_t)i * F + c] = g[c];
    }
}

// phase 4 (backward)
__global__ void __launch_bounds__(256)
contrast_dfeat_kernel(int N, int F, int K, const float* __restrict__ g, const int* __restrict__ labels,
                      const float* __restrict__ dU, const float* __restrict__ counts,
                      const float* __restrict__ inv_norm, bool means, const float* __restrict__ grad_scale,
                      float* __restrict__ dfeat) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * F) return;
    const int i = e / F, c = e - i * F;
    const int y = labels[i];
    float v = g[e];
    if (means && y >= 0 && y < K) v = fmaf(dU[(size_t)y * F + c], 1.0f / counts[y], v);
    dfeat[e] = v * (grad_scale ? *grad_scale : 1.0f) * inv_norm[i];
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
    if (N <= 0 || K <= 0) {
        ISR_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), stream));
        return ISR_OK;
    }
    float* fhat = reinterpret_cast<float*>(w + L.fhat);
    float* inv_norm = reinterpret_cast<float*>(w + L.inv_norm);
    float* sums = reinterpret_cast<float*>(w + L.sums);
    float* counts = reinterpret_cast<float*>(w + L.counts);
    float* spread = reinterpret_cast<float*>(w + L.spread);
    ISR_CUDA_TRY(cudaMemsetAsync(w + L.zeroed, 0, L.zeroed_bytes, stream));
    const int blocks = (N + kCB - 1) / kCB, FS = sample_stride(F);
    const size_t smem1 = ((size_t)kCB * FS + kCB) * 4;
    const size_t smem2 = ((size_t)K * F + 2 * kCB) * 4;
    const size_t smem3 = ((size_t)K * F + ((K + 3) & ~3) + (size_t)kCB * FS + (size_t)kKC * kCB) * 4;
    auto run = [&](auto stats_k, auto loss_k) -> int {
        ISR_CUDA_TRY(cudaFuncSetAttribute(stats_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        ISR_CUDA_TRY(cudaFuncSetAttribute(contrast_spread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        ISR_CUDA_TRY(cudaFuncSetAttribute(loss_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        stats_k<<<blocks, kCB, smem1, stream>>>(N, F, K, features, labels, predef_u == nullptr, fhat, inv_norm, sums, counts,
                                                loss);
        contrast_spread_kernel<<<blocks, kCB, smem2, stream>>>(N, F, K, fhat, labels, predef_u, sums, counts, spread);
        loss_k<<<blocks, kCB, smem3, stream>>>(N, F, K, fhat, labels, predef_u, sums, counts, spread, temp_lambda,
                                               reinterpret_cast<float*>(w + L.g), reinterpret_cast<float*>(w + L.dU), loss);
        ISR_CUDA_TRY(cudaGetLastError());
        return ISR_OK;
    };
    if (F <= 4) return run(contrast_stats_kernel<4>, contrast_loss_kernel<4>);
    if (F <= 8) return run(contrast_stats_kernel<8>, contrast_loss_kernel<8>);
    if (F <= 16) return run(contrast_stats_kernel<16>, contrast_loss_kernel<16>);
    if (F <= 24) return run(contrast_stats_kernel<24>, contrast_loss_kernel<24>);
    return run(contrast_stats_kernel<32>, contrast_loss_kernel<32>);
}

int launch_contrastive_bwd(int N, int F, int K, const int* labels, const float* predef_u, const void* ws,
                           const float* grad_scale, float* dfeat, cudaStream_t stream) {
    if (N <= 0) return ISR_OK;
    ContrastWs L(N, F, K);
    const char* w = static_cast<const char*>(ws);
    const int total = N * F;
    contrast_dfeat_kernel<<<(total + 255) / 256, 256, 0, stream>>>(
        N, F, K, reinterpret_cast<const float*>(w + L.g), labels, reinterpret_cast<const float*>(w + L.dU),
        reinterpret_cast<const float*>(w + L.counts), reinterpret_cast<const float*>(w + L.inv_norm),
        predef_u == nullptr, grad_scale, dfeat);
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

}  // namespace isr
