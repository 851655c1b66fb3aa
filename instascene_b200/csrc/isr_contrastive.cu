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
// The [N,F]x[F,K] logits contraction (~70 MFLOP at N=32768,K=64,F=16) is the one GEMM-shaped piece of the path: it runs
// on the tensor cores (tcgen05.mma kind::tf32 with a 3xTF32 split, fp32 accumulation in TMEM), see phase 3 below.
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
// phase one kernel, one sample per thread (256-thread blocks; the tensor-core loss kernel: 128 = the rows of a UMMA tile):
//   stats   f^ = f/(|f|+1e-9); per-cluster sums and counts
//   spread  u_k = mean (or predefined prototype); per-cluster sum of |f^ - u_k|
//   loss    phi_k; logits (tcgen05), loss, softmax coefficients coef_ik = (p_ik - [k == y_i]) / phi_k (never stored: a block
//           keeps 64 clusters x 128 samples of them in shared memory), g_i = sum_k coef_ik u_k, dU_k = sum_i coef_ik f^_i
//   dfeat   (backward) dL/df_i = grad_scale / (|f_i| + eps) * (g_i + [means] dU[y_i] / n_{y_i})
// Per-cluster block reductions are "owner computes": the block's samples and labels sit in shared memory and a thread
// sums the members of the (cluster, channel) entries it owns -- no shared-memory float atomics (CAS loops on this
// target); one global red per entry per block.
constexpr int kCB = 256;   // samples per block
constexpr int kKC = 64;    // clusters per coefficient chunk

// sum over the block (kBlock threads); s_red must hold kBlock/32 floats
template <int kBlock = kCB>
__device__ __forceinline__ float block_sum(float v, float* s_red) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    for (int w = 0; w < kBlock / 32; w++) t += s_red[w];
    return t;
}

// block partial of  out[k][c] += sum_{i in block, label_i == k} w_i * sf[i][c]   for k in [k0, k0+nk)
// (w == nullptr: w_i = 1).  sw is indexed [k - k0][i] when per-cluster weights are given (kPerCluster).
template <bool kPerCluster, int kBlock = kCB>
__device__ __forceinline__ void owner_accumulate(int F, int FS, int k0, int nk, const float* __restrict__ sf,
                                                 const int* __restrict__ sl, const float* __restrict__ sw,
                                                 float* __restrict__ out) {
    if ((F & 3) == 0) {
        const int F4 = F >> 2;
        for (int e = threadIdx.x; e < nk * F4; e += kBlock) {
            const int kk = e / F4, c4 = e - kk * F4, k = k0 + kk;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = 0; i < kBlock; i++) {
                float w;
                if (kPerCluster) { w = sw[kk * kBlock + i]; if (w == 0.0f) continue; }
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
        for (int e = threadIdx.x; e < nk * F; e += kBlock) {
            const int kk = e / F, c = e - kk * F, k = k0 + kk;
            float acc = 0.0f;
            for (int i = 0; i < kBlock; i++) {
                float w;
                if (kPerCluster) { w = sw[kk * kBlock + i]; if (w == 0.0f) continue; }
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
    // owner computes, label driven: the first sample of the block that carries a label owns that cluster for this block
    // (cost independent of K): count its members and, for cluster means, sum their rows -- one global red per value
    int* sfirst = sl + kCB;  // [kCB] 1 = first occurrence of its label in the block
    {
        bool first = y >= 0;
        if (first)
            for (int j = 0; j < (int)threadIdx.x; j++) if (sl[j] == y) { first = false; break; }
        sfirst[threadIdx.x] = first ? 1 : 0;
        if (first) {
            int n = 0;
            for (int j = threadIdx.x; j < kCB; j++) n += (sl[j] == y);
            atomicAdd(counts + y, (float)n);
        }
    }
    if (want_sums) {
        __syncthreads();
        // entries (owner sample, group of 4 channels) spread over the threads; members summed in ascending sample order
        const int G = (F + 3) >> 2;
        for (int e = threadIdx.x; e < kCB * G; e += kCB) {
            const int j0 = e / G, cg = e - j0 * G;
            if (!sfirst[j0]) continue;
            const int k = sl[j0];
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = j0; j < kCB; j++) {
                if (sl[j] != k) continue;
#pragma unroll
                for (int c = 0; c < 4; c++) if (4 * cg + c < F) acc[c] += sf[j * FS + 4 * cg + c];
            }
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (4 * cg + c < F && acc[c] != 0.0f) atomicAdd(sums + (size_t)k * F + 4 * cg + c, acc[c]);
        }
    }
}

// phase 2   (a sample needs only ITS cluster's centre: read from global memory, so K is not bounded by shared memory)
__global__ void __launch_bounds__(kCB)
contrast_spread_kernel(int N, int F, int K, const float* __restrict__ fhat, const int* __restrict__ labels,
                       const float* __restrict__ predef_u, const float* __restrict__ sums,
                       const float* __restrict__ counts, float* __restrict__ spread) {
    __shared__ float sd[kCB];
    __shared__ int sl[kCB];
    const int i = blockIdx.x * kCB + threadIdx.x;
    int y = -1;
    float d = 0.0f;
    if (i < N) {
        y = labels[i];
        if (y < 0 || y >= K) y = -1;
        if (y >= 0) {
            const float n = counts[y];
            float ss = 0.0f;
            for (int c = 0; c < F; c++) {
                const float u = predef_u ? predef_u[(size_t)y * F + c] : (n > 0.0f ? sums[(size_t)y * F + c] / n : 0.0f);
                const float t = fhat[(size_t)i * F + c] - u;
                ss = fmaf(t, t, ss);
            }
            d = sqrtf(ss);
        }
    }
    sd[threadIdx.x] = d;
    sl[threadIdx.x] = y;
    __syncthreads();
    // owner computes: the clusters present in this block are few; walk the block's labels once per distinct label
    // (first occurrence owns it) instead of once per cluster id, so the cost does not grow with K
    if (y >= 0) {
        bool first = true;
        for (int j = 0; j < (int)threadIdx.x; j++) if (sl[j] == y) { first = false; break; }
        if (first) {
            float sacc = 0.0f;
            for (int j = threadIdx.x; j < kCB; j++) if (sl[j] == y) sacc += sd[j];
            if (sacc != 0.0f) atomicAdd(spread + y, sacc);
        }
    }
}

// phase 3 -- the [N,F] x [F,K] similarity matrix on the 5th-generation tensor cores (tcgen05, accumulators in TMEM).
//
// One CTA = 128 samples = the M = 128 rows of a UMMA tile; the K cluster centres (pre-scaled by 1/phi_k) are the N
// columns, the feature dim F is the contraction.  Operands are staged in shared memory in the canonical K-major
// no-swizzle layout (8-row x 16-byte core matrices; a "panel" = the same 4 features of all rows, 16 bytes per row).
// fp32 fidelity: each operand is split x = hi + lo with hi = the 19 leading bits (exactly a TF32 number), and
// hi*hi + lo*hi + hi*lo is accumulated in fp32 in TMEM (3xTF32: the dropped lo*lo and the 11-bit read of lo are
// ~2^-22 relative).  One elected thread issues the 3*F/8 tcgen05.mma (kind::tf32, M128 x N<=256 x K8) per column
// chunk and commits them to an mbarrier; every thread then reads ITS sample's row of logits straight from TMEM
// (tcgen05.ld 32x32b: lane = row) -- pass 0 for the softmax denominator and the loss, pass 1 for the coefficients
// coef_ik = (p_ik - [k == y_i]) / phi_k, from which g_i = sum_k coef_ik u_k and dU_k = sum_i coef_ik f^_i are formed
// as before.  With more than 256 clusters the columns are processed in chunks of 256 and pass 1 re-issues the MMAs.
constexpr int kLB = 128;  // samples per CTA of the tensor-core loss kernel

namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start address, leading (K-direction) and stride (row-group) byte offsets in 16-byte units, version 1 (sm_100).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
__device__ __forceinline__ uint32_t instr_desc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}\n"
        :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(cols) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + lane id)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
}  // namespace tc

struct ContrastTcSmem {  // dynamic shared-memory carve-up of contrast_loss_tc_kernel (byte offsets)
    // Everything per COLUMN CHUNK of at most 256 clusters (the N extent of one UMMA tile): the shared-memory footprint
    // does not depend on K, chunks beyond the first are re-staged from global memory (any K is supported).
    int Kpad, NC, tmem_cols;
    size_t su, sphi, sf, sw, apan, bpan, total;
    __host__ __device__ ContrastTcSmem(int F, int FP, int K) {
        Kpad = (K + 31) & ~31;
        NC = Kpad < 256 ? Kpad : 256;
        tmem_cols = 32;
        while (tmem_cols < NC) tmem_cols <<= 1;
        const int FS = (F + 3) & ~3;
        size_t o = 0;
        su = o;   o += (size_t)NC * F * 4;
        o = (o + 15) & ~(size_t)15;
        sphi = o; o += (size_t)NC * 4;
        sf = o;   o += (size_t)kLB * FS * 4;
        sw = o;   o += (size_t)kKC * kLB * 4;
        o = (o + 127) & ~(size_t)127;
        apan = o; o += (size_t)2 * (FP / 4) * kLB * 16;    // [hi|lo][FP/4 panels][128 rows] x 16 B
        bpan = o; o += (size_t)2 * (FP / 4) * NC * 16;     // [hi|lo][FP/4 panels][NC rows] x 16 B
        total = o;
    }
};

template <int FP>  // F padded to a multiple of 8 (one tcgen05.mma contracts 8 TF32 values)
__global__ void __launch_bounds__(kLB)
contrast_loss_tc_kernel(int N, int F, int K, const float* __restrict__ fhat, const int* __restrict__ labels,
                        const float* __restrict__ predef_u, const float* __restrict__ sums,
                        const float* __restrict__ counts, const float* __restrict__ spread, float temp_lambda,
                        float min_count, float* __restrict__ g_out, float* __restrict__ dU, float* __restrict__ loss) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_red[kLB / 32];
    const ContrastTcSmem L(F, FP, K);
    const int FS = sample_stride(F), Kpad = L.Kpad, NC = L.NC;
    float* su = reinterpret_cast<float*>(smem_raw + L.su);      // [NC][F]   centres of the current column chunk (for g)
    float* sphi = reinterpret_cast<float*>(smem_raw + L.sphi);  // [NC]      1/phi_k, 0 for an absent / dropped / padding cluster
    float* sf = reinterpret_cast<float*>(smem_raw + L.sf);      // [128][FS] this CTA's samples (for dU)
    float* sw = reinterpret_cast<float*>(smem_raw + L.sw);      // [kKC][128] coefficient chunk (for dU)
    float4* apan = reinterpret_cast<float4*>(smem_raw + L.apan);
    float4* bpan = reinterpret_cast<float4*>(smem_raw + L.bpan);
    constexpr int NP = FP / 4;  // 16-byte K panels
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool means = predef_u == nullptr;

    // Stages column chunk [n0, n0 + NC): centres, 1/phi (0 = the cluster does not take part: absent, or not more than
    // min_count members -- utils/contrastive_utils.py:33-35), and the B operand panels (centres scaled by 1/phi so that
    // the MMA yields the logits; zero rows for clusters that do not take part).
    auto stage_chunk = [&](int n0) {
        for (int e = tid; e < NC * F; e += kLB) {
            const int k = n0 + e / F;
            float u = 0.0f;
            if (k < K) {
                const float n = counts[k];
                u = predef_u ? predef_u[(size_t)n0 * F + e] : (n > 0.0f ? sums[(size_t)n0 * F + e] / n : 0.0f);
            }
            su[e] = u;
        }
        for (int kk = tid; kk < NC; kk += kLB) {
            const int k = n0 + kk;
            float ip = 0.0f;
            if (k < K) {
                const float n = counts[k];
                if (n > min_count) ip = 1.0f / fminf(fmaxf(10.0f * (spread[k] / (n * logf(n + temp_lambda))), 0.5f), 1.0f);
            }
            sphi[kk] = ip;
        }
        __syncthreads();
        for (int e = tid; e < NP * NC; e += kLB) {
            const int p = e / NC, n = e - p * NC;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = 4 * p + j;
                v[j] = (c < F) ? su[(size_t)n * F + c] * sphi[n] : 0.0f;
            }
            const float4 h = make_float4(tc::tf32_hi(v[0]), tc::tf32_hi(v[1]), tc::tf32_hi(v[2]), tc::tf32_hi(v[3]));
            bpan[p * NC + n] = h;
            bpan[(NP + p) * NC + n] = make_float4(v[0] - h.x, v[1] - h.y, v[2] - h.z, v[3] - h.w);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // operand panels -> visible to the tensor-core (async) proxy
        __syncthreads();
    };

    const int i = blockIdx.x * kLB + tid;
    int y = -1;
    {   // A operand: this thread's sample row, split hi / lo, one float4 per panel
        float f[FP];
#pragma unroll
        for (int c = 0; c < FP; c++) f[c] = 0.0f;
        if (i < N) {
            y = labels[i];
            if (y < 0 || y >= K) y = -1;
            if (y >= 0 && !(counts[y] > min_count)) y = -1;  // its cluster was dropped
#pragma unroll
            for (int c = 0; c < FP; c++)
                if (c < F) { f[c] = fhat[(size_t)i * F + c]; sf[tid * FS + c] = f[c]; }
        } else {
            for (int c = 0; c < F; c++) sf[tid * FS + c] = 0.0f;
        }
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const float4 h = make_float4(tc::tf32_hi(f[4 * p]), tc::tf32_hi(f[4 * p + 1]), tc::tf32_hi(f[4 * p + 2]), tc::tf32_hi(f[4 * p + 3]));
            apan[p * kLB + tid] = h;
            apan[(NP + p) * kLB + tid] = make_float4(f[4 * p] - h.x, f[4 * p + 1] - h.y, f[4 * p + 2] - h.z, f[4 * p + 3] - h.w);
        }
    }
    const uint32_t mbar = tc::smem_u32(&s_mbar);
    if (tid == 0) {
        tc::mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(&s_tmem), (uint32_t)L.tmem_cols);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t a_base = tc::smem_u32(apan), b_base = tc::smem_u32(bpan);
    const int nch = (Kpad + NC - 1) / NC;
    uint32_t uses = 0;

    const bool valid = y >= 0;
    float sum = 0.0f, dy = 0.0f, inv_denom = 0.0f;
    float g[FP];
#pragma unroll
    for (int c = 0; c < FP; c++) g[c] = 0.0f;

    for (int pass = 0; pass < 2; pass++) {
        for (int ch = 0; ch < nch; ch++) {
            const int n0 = ch * NC, ncols = min(NC, Kpad - n0);
            if (pass == 0 || nch > 1) {
                stage_chunk(n0);
                if (tid == 0) {
                    const uint32_t idesc = tc::instr_desc_tf32(kLB, ncols);
                    uint32_t acc = 0;
#pragma unroll
                    for (int ks = 0; ks < FP / 8; ks++) {
#pragma unroll
                        for (int t = 0; t < 3; t++) {  // hi*hi, lo*hi, hi*lo
                            const int pa = (t == 1 ? NP : 0) + 2 * ks, pb = (t == 2 ? NP : 0) + 2 * ks;
                            const uint64_t da = tc::smem_desc(a_base + (uint32_t)pa * kLB * 16, kLB * 16, 128);
                            const uint64_t db = tc::smem_desc(b_base + (uint32_t)(pb * NC) * 16, (uint32_t)NC * 16, 128);
                            tc::mma_tf32(tmem, da, db, idesc, acc);
                            acc = 1;
                        }
                    }
                    tc::commit(mbar);
                }
                tc::mbar_wait(mbar, uses & 1u);
                uses++;
                __syncwarp();  // lane 0 issued the MMAs: reconverge before the .aligned TMEM loads
                tc::fence_after();
            }
            // this thread's row: lane 32*warp + lane id, columns [0, ncols)
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                float lg[32];
                tc::tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, lg);
                if (pass == 0) {
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const int kk = c0 + j;
                            if (sphi[kk] != 0.0f) {
                                const float e = expf(lg[j]);
                                sum += e;
                                if (n0 + kk == y) dy = e;
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int kk = c0 + j;
                        const float ip = sphi[kk];
                        float cf = 0.0f;
                        if (valid && ip != 0.0f) {
                            cf = (expf(lg[j]) * inv_denom - (n0 + kk == y ? 1.0f : 0.0f)) * ip;
#pragma unroll
                            for (int c = 0; c < FP; c++)
                                if (c < F) g[c] = fmaf(cf, su[(size_t)kk * F + c], g[c]);
                        }
                        if (means) sw[((c0 + j) & (kKC - 1)) * kLB + tid] = cf;
                    }
                    if (means && (((c0 + 32) & (kKC - 1)) == 0 || c0 + 32 >= ncols)) {  // a kKC-cluster chunk of coefficients is complete
                        const int k0 = n0 + ((c0 + 32 - 1) & ~(kKC - 1));
                        const int nk = min(n0 + c0 + 32, K) - k0;
                        __syncthreads();
                        if (nk > 0) owner_accumulate<true, kLB>(F, FS, k0, nk, sf, nullptr, sw, dU);
                        __syncthreads();
                    }
                }
            }
            if (pass == 0 || nch > 1) {  // TMEM and the chunk's shared-memory panels are overwritten by the next chunk
                tc::fence_before();
                __syncthreads();
                tc::fence_after();
            }
        }
        if (pass == 0) {
            const float denom = sum + 1e-9f;
            const float li = block_sum<kLB>(valid ? -logf(dy / denom) : 0.0f, s_red);
            if (tid == 0 && li != 0.0f) atomicAdd(loss, li);
            inv_denom = 1.0f / denom;
        }
    }
    if (i < N) {
#pragma unroll
        for (int c = 0; c < FP; c++)
            if (c < F) g_out[(size_t)i * F + c] = g[c];
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)L.tmem_cols);
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
    if (fwd) { rownorm_fwd_kernel<FP><<<(P + 255) / 256, 256, 0, stream>>>(P, F, x, e1, e2, stages, out); note_launch(); }
    else { rownorm_bwd_kernel<FP><<<(P + 255) / 256, 256, 0, stream>>>(P, F, x, dy, e1, e2, stages, out); note_launch(); }
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
// Bias corrections: from the host step count, or -- step_dev != nullptr -- from a DEVICE step counter, so that a captured
// CUDA graph replays with the right correction (one thread per block evaluates the two pow() in double, as the host does).
struct AdamCoef { float step_size, inv_sqrt_bias2; };
__device__ __forceinline__ AdamCoef adam_coef(float lr, float beta1, float beta2, float step_size_host, float isb2_host,
                                              const int* __restrict__ step_dev) {
    __shared__ AdamCoef s_coef;
    if (step_dev == nullptr) return AdamCoef{step_size_host, isb2_host};
    if (threadIdx.x == 0) {
        const double st = (double)max(1, *step_dev);
        s_coef.step_size = (float)((double)lr / (1.0 - pow((double)beta1, st)));
        s_coef.inv_sqrt_bias2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, st)));
    }
    __syncthreads();
    return s_coef;
}

__global__ void __launch_bounds__(256)
adam_step_kernel(size_t n4, size_t n, float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                 float4* __restrict__ v, float lr, float beta1, float beta2, float eps, float step_size_host, float isb2_host,
                 const int* __restrict__ step_dev) {
    const AdamCoef co = adam_coef(lr, beta1, beta2, step_size_host, isb2_host, step_dev);
    const float step_size = co.step_size, inv_sqrt_bias2 = co.inv_sqrt_bias2;
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

// Adam on a [P,F] parameter whose gradient arrives as dL/d(normalised rows): the backward of the one / two row
// normalisations (rownorm_bwd_kernel) is applied on the fly, so x, dy, exp_avg and exp_avg_sq are each read once and the
// chained gradient is never written (rownorm_bwd + adam_step: 3 + 7 passes over [P,F]; here 7).  g_extra: an ordinary
// gradient of the same parameter from other uses (added after the chain rule), or nullptr.
template <int FP>
__global__ void __launch_bounds__(256)
adam_rownorm_kernel(int P, int F, float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ g_extra,
                    float* __restrict__ m, float* __restrict__ v, float e1, float e2, int stages, float lr, float beta1,
                    float beta2, float eps, float step_size_host, float isb2_host, const int* __restrict__ step_dev) {
    const AdamCoef co = adam_coef(lr, beta1, beta2, step_size_host, isb2_host, step_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float xv[FP], d[FP];
    const bool vec = (F & 3) == 0;
    auto load = [&](const float* src, float* dst) {
#pragma unroll
        for (int c = 0; c < FP; c += 4) {
            if (vec && c < F) {
                const float4 t = *reinterpret_cast<const float4*>(src + (size_t)i * F + c);
                dst[c] = t.x; dst[c + 1] = t.y; dst[c + 2] = t.z; dst[c + 3] = t.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) dst[c + k] = (c + k < F) ? src[(size_t)i * F + c + k] : 0.0f;
            }
        }
    };
    auto store = [&](float* dst, const float* src) {
#pragma unroll
        for (int c = 0; c < FP; c += 4) {
            if (vec && c < F) {
                *reinterpret_cast<float4*>(dst + (size_t)i * F + c) = make_float4(src[c], src[c + 1], src[c + 2], src[c + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) if (c + k < F) dst[(size_t)i * F + c + k] = src[c + k];
            }
        }
    };
    load(x, xv);
    load(dy, d);
    // chain rule of the normalisations, innermost last (same arithmetic as rownorm_bwd_kernel)
    if (stages > 1) {
        float y[FP], ss = 0.0f;
#pragma unroll
        for (int c = 0; c < FP; c++) ss = fmaf(xv[c], xv[c], ss);
        const float inv = 1.0f / (sqrtf(ss) + e1);
#pragma unroll
        for (int c = 0; c < FP; c++) y[c] = xv[c] * inv;
        rownorm_vjp<FP>(y, d, e2);
    }
    rownorm_vjp<FP>(xv, d, e1);
    if (g_extra != nullptr) {
        float ge[FP];
        load(g_extra, ge);
#pragma unroll
        for (int c = 0; c < FP; c++) d[c] += ge[c];
    }
    float mv[FP], vv[FP];
    load(m, mv);
    load(v, vv);
#pragma unroll
    for (int c = 0; c < FP; c++) {
        mv[c] = beta1 * mv[c] + (1.0f - beta1) * d[c];
        vv[c] = beta2 * vv[c] + (1.0f - beta2) * d[c] * d[c];
        xv[c] -= co.step_size * mv[c] / (sqrtf(vv[c]) * co.inv_sqrt_bias2 + eps);
    }
    store(x, xv);
    store(m, mv);
    store(v, vv);
}

static void adam_host_coef(float lr, float beta1, float beta2, int step, float& step_size, float& isb2) {
    const int st = step > 0 ? step : 1;
    const double b1 = 1.0 - pow((double)beta1, (double)st), b2 = 1.0 - pow((double)beta2, (double)st);
    step_size = (float)((double)lr / b1);
    isb2 = (float)(1.0 / sqrt(b2));
}

int launch_adam(size_t n, float* p, const float* g, float* m, float* v, float lr, float beta1, float beta2, float eps,
                int step, const int* step_dev, cudaStream_t stream) {
    if (n == 0) return ISR_OK;
    float step_size, isb2;
    adam_host_coef(lr, beta1, beta2, step, step_size, isb2);
    const size_t n4 = n / 4;
    adam_step_kernel<<<(unsigned)((n4 + 1 + 255) / 256), 256, 0, stream>>>(
        n4, n, reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
        reinterpret_cast<float4*>(v), lr, beta1, beta2, eps, step_size, isb2, step_dev); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

// Same update for F % 4 == 0 with LPR lanes per row, one float4 per lane (LPR = 1, 2, 4, 8 for F <= 4, 8, 16, 32): every
// load / store instruction of a warp covers whole 128-byte lines (one thread per row reads 16 of every 64..128 bytes per
// instruction and keeps 5 F-float arrays in registers), the row sums are two shuffle-reduction steps.
template <int LPR>
__global__ void __launch_bounds__(256)
adam_rownorm_vec_kernel(int P, int F, float4* __restrict__ x, const float4* __restrict__ dy, const float4* __restrict__ g_extra,
                        float4* __restrict__ m, float4* __restrict__ v, float e1, float e2, int stages, float lr, float beta1,
                        float beta2, float eps, float step_size_host, float isb2_host, const int* __restrict__ step_dev) {
    const AdamCoef co = adam_coef(lr, beta1, beta2, step_size_host, isb2_host, step_dev);
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t row = t / LPR;
    const int sub = (int)(t % LPR), f4 = F >> 2;
    const bool on = row < (size_t)P && sub < f4;
    const size_t idx = row * (size_t)f4 + sub;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 X = z4, D = z4, M = z4, V = z4, G = z4;
    if (on) {
        X = x[idx]; D = dy[idx]; M = m[idx]; V = v[idx];
        if (g_extra != nullptr) G = g_extra[idx];
    }
    auto row_sum = [&](float a) {
#pragma unroll
        for (int o = LPR >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        return a;
    };
    auto dot4 = [](const float4 a, const float4 b) { return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x))); };
    // d <- vjp of y = u / (|u| + e) at u, cotangent d  (rownorm_vjp)
    auto vjp = [&](const float4 u, float4& d, float e) {
        const float ss = row_sum(dot4(u, u)), ud = row_sum(dot4(u, d));
        const float n = sqrtf(ss), ne = n + e;
        const float a = 1.0f / ne;
        const float b = n > 0.0f ? ud / (n * ne * ne) : 0.0f;
        d.x = fmaf(-b, u.x, a * d.x); d.y = fmaf(-b, u.y, a * d.y); d.z = fmaf(-b, u.z, a * d.z); d.w = fmaf(-b, u.w, a * d.w);
    };
    if (stages > 1) {
        const float inv = 1.0f / (sqrtf(row_sum(dot4(X, X))) + e1);
        const float4 Y = make_float4(X.x * inv, X.y * inv, X.z * inv, X.w * inv);
        vjp(Y, D, e2);
    }
    vjp(X, D, e1);
    D.x += G.x; D.y += G.y; D.z += G.z; D.w += G.w;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        mm = beta1 * mm + (1.0f - beta1) * gg;
        vv = beta2 * vv + (1.0f - beta2) * gg * gg;
        pp -= co.step_size * mm / (sqrtf(vv) * co.inv_sqrt_bias2 + eps);
    };
    upd(X.x, D.x, M.x, V.x); upd(X.y, D.y, M.y, V.y); upd(X.z, D.z, M.z, V.z); upd(X.w, D.w, M.w, V.w);
    if (on) { x[idx] = X; m[idx] = M; v[idx] = V; }
}

template <int FP>
static int adam_rownorm_launch(int P, int F, float* x, const float* dy, const float* g_extra, float* m, float* v, float e1,
                               float e2, int stages, float lr, float beta1, float beta2, float eps, int step,
                               const int* step_dev, cudaStream_t stream) {
    float step_size, isb2;
    adam_host_coef(lr, beta1, beta2, step, step_size, isb2);
    if ((F & 3) == 0) {  // (16-byte alignment of the arrays is checked at the boundary)
        constexpr int LPR = FP <= 4 ? 1 : (FP <= 8 ? 2 : (FP <= 16 ? 4 : 8));
        const size_t threads = (size_t)P * LPR;
        adam_rownorm_vec_kernel<LPR><<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
            P, F, reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(g_extra),
            reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), e1, e2, stages, lr, beta1, beta2, eps, step_size, isb2,
            step_dev); note_launch();
        ISR_CUDA_TRY(cudaGetLastError());
        return ISR_OK;
    }
    adam_rownorm_kernel<FP><<<(P + 255) / 256, 256, 0, stream>>>(P, F, x, dy, g_extra, m, v, e1, e2, stages, lr, beta1, beta2,
                                                                 eps, step_size, isb2, step_dev); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_adam_rownorm(int P, int F, float* x, const float* dy, const float* g_extra, float* m, float* v, float e1, float e2,
                        int stages, float lr, float beta1, float beta2, float eps, int step, const int* step_dev,
                        cudaStream_t stream) {
    if (P <= 0 || F <= 0) return ISR_OK;
#define ISR_AR(FPV) return adam_rownorm_launch<FPV>(P, F, x, dy, g_extra, m, v, e1, e2, stages, lr, beta1, beta2, eps, step, step_dev, stream)
    if (F <= 4) ISR_AR(4);
    if (F <= 8) ISR_AR(8);
    if (F <= 16) ISR_AR(16);
    if (F <= 24) ISR_AR(24);
    ISR_AR(32);
#undef ISR_AR
}

size_t contrastive_ws_bytes(int N, int F, int K) { return ContrastWs(N, F, K).total; }

int launch_gather_pixels(int F, int64_t HW, const float* map, int n, const int* pix_ids, float* out, cudaStream_t stream) {
    if (n <= 0 || F <= 0) return ISR_OK;
    const int total = n * F;
    gather_pixels_kernel<<<(total + 255) / 256, 256, 0, stream>>>(F, HW, map, n, pix_ids, out); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_contrastive_fwd(int N, int F, int K, const float* features, const int* labels, const float* predef_u,
                           float temp_lambda, int min_pixnum, void* ws, float* loss, cudaStream_t stream) {
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
    const size_t smem1 = ((size_t)kCB * FS + 2 * kCB) * 4;
    auto run = [&](auto stats_k, auto loss_k, int FPT) -> int {
        const ContrastTcSmem S(F, FPT, K);
        const size_t smem3 = S.total;
        if (smem3 > 226 * 1024) return ISR_ERR_UNSUPPORTED;  // (177 KB at F = 32: cannot happen for F <= 32)
        ISR_CUDA_TRY(cudaFuncSetAttribute(stats_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        ISR_CUDA_TRY(cudaFuncSetAttribute(loss_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        stats_k<<<blocks, kCB, smem1, stream>>>(N, F, K, features, labels, predef_u == nullptr, fhat, inv_norm, sums, counts,
                                                loss); note_launch();
        contrast_spread_kernel<<<blocks, kCB, 0, stream>>>(N, F, K, fhat, labels, predef_u, sums, counts, spread); note_launch();
        loss_k<<<(N + kLB - 1) / kLB, kLB, smem3, stream>>>(N, F, K, fhat, labels, predef_u, sums, counts, spread, temp_lambda,
                                                            (float)(min_pixnum > 0 ? min_pixnum : 0),
                                                            reinterpret_cast<float*>(w + L.g), reinterpret_cast<float*>(w + L.dU),
                                                            loss); note_launch();
        ISR_CUDA_TRY(cudaGetLastError());
        return ISR_OK;
    };
    if (F <= 4) return run(contrast_stats_kernel<4>, contrast_loss_tc_kernel<8>, 8);
    if (F <= 8) return run(contrast_stats_kernel<8>, contrast_loss_tc_kernel<8>, 8);
    if (F <= 16) return run(contrast_stats_kernel<16>, contrast_loss_tc_kernel<16>, 16);
    if (F <= 24) return run(contrast_stats_kernel<24>, contrast_loss_tc_kernel<24>, 24);
    return run(contrast_stats_kernel<32>, contrast_loss_tc_kernel<32>, 32);
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
        predef_u == nullptr, grad_scale, dfeat); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

}  // namespace isr
