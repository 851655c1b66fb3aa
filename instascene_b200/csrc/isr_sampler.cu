// isr_sampler.cu -- "n labelled pixels, uniformly, with replacement" in two kernels.
// Reference: train_semantic.py:118-129 -- `valid = segmap > 0; idx = torch.randint(0, len(valid), (n,))` after a
// boolean-mask gather of the whole [F,H,W] feature map.  Round 1 drew the same samples without that gather but with
// ~10 torch launches (mask, cumsum over H*W, rand, mul, cast, clamp, add, searchsorted, clamp, index).  Here:
//   label_words_kernel   one pass over the label map: a 32-pixel occupancy word per warp ballot + the number of labelled
//                        pixels per 4096-pixel block
//   label_select_kernel  per sample: rank t = min(int(u * n_valid), n_valid - 1) (fp32 product, truncation: the value
//                        round 1's torch expression produced), binary search over the block prefix (built per CTA in
//                        shared memory), popcount walk over the block's 128 words, __fns for the bit; writes the pixel
//                        id and its label
#include "isr_common.cuh"

namespace isr {

constexpr int kSampBlockPix = 4096;
constexpr int kSampWords = kSampBlockPix / 32;

struct SamplerWs {
    size_t words, counts, total;
    int nblocks;
    explicit SamplerWs(int64_t HW) {
        nblocks = (int)((HW + kSampBlockPix - 1) / kSampBlockPix);
        size_t o = 0;
        words = o;  o = align_up(o + (size_t)(nblocks > 0 ? nblocks : 1) * kSampWords * 4, 256);
        counts = o; o = align_up(o + (size_t)(nblocks > 0 ? nblocks : 1) * 4, 256);
        total = o;
    }
};

template <typename T>
__global__ void __launch_bounds__(256) label_words_kernel(int64_t HW, const T* __restrict__ labels, uint32_t* __restrict__ words,
                                                         uint32_t* __restrict__ block_counts) {
    __shared__ uint32_t s_cnt[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * kSampBlockPix;
    uint32_t cnt = 0;
#pragma unroll 4
    for (int it = 0; it < kSampBlockPix / 256; it++) {
        const int64_t p = base + it * 256 + threadIdx.x;
        const bool on = p < HW && labels[p] > 0;
        const unsigned w = __ballot_sync(0xffffffffu, on);
        if (lane == 0) {
            words[(size_t)blockIdx.x * kSampWords + it * 8 + warp] = w;
            cnt += (uint32_t)__popc(w);
        }
    }
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; w++) t += s_cnt[w];
        block_counts[blockIdx.x] = t;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) label_select_kernel(int64_t HW, int n, const float* __restrict__ u,
                                                          const uint32_t* __restrict__ words,
                                                          const uint32_t* __restrict__ block_counts, int nblocks,
                                                          const T* __restrict__ labels, int64_t* __restrict__ pix_out,
                                                          int* __restrict__ lab_out) {
    extern __shared__ uint32_t prefix[];  // [nblocks + 1] exclusive scan of the block counts
    __shared__ uint32_t s_part[256];
    const int tid = threadIdx.x;
    const int per = (nblocks + 255) / 256, b0 = tid * per, b1 = min(nblocks, b0 + per);
    uint32_t mine = 0;
    for (int b = b0; b < b1; b++) mine += block_counts[b];
    s_part[tid] = mine;
    __syncthreads();
    uint32_t run = 0;
    for (int t = 0; t < tid; t++) run += s_part[t];   // 256 partials: a serial prefix per thread is ~130 adds on average
    for (int b = b0; b < b1; b++) { prefix[b] = run; run += block_counts[b]; }
    if (tid == 255) prefix[nblocks] = run;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + tid;
    if (i >= n) return;
    const uint32_t total = prefix[nblocks];
    if (total == 0u) {  // nothing labelled (the reference's randint(0, 0) raises): pixel 0 with its (<= 0) label
        pix_out[i] = 0;
        lab_out[i] = HW > 0 ? (int)labels[0] : 0;
        return;
    }
    // torch: clamp((u * n_valid).to(int32), max = n_valid - 1): fp32 product, truncation towards zero
    uint32_t t = (uint32_t)max(0, __float2int_rz(__fmul_rn(u[i], (float)total)));
    if (t > total - 1u) t = total - 1u;
    int lo = 0, hi = nblocks;  // prefix[lo] <= t < prefix[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (prefix[mid] <= t) lo = mid; else hi = mid;
    }
    uint32_t r = t - prefix[lo];
    const uint32_t* wp = words + (size_t)lo * kSampWords;
    int w = 0;
    uint32_t word = __ldg(wp);
    for (;;) {
        const uint32_t c = (uint32_t)__popc(word);
        if (r < c || w == kSampWords - 1) break;
        r -= c;
        word = __ldg(wp + ++w);
    }
    const int bit = (int)__fns(word, 0u, (int)r + 1);
    const int64_t pix = (int64_t)lo * kSampBlockPix + w * 32 + (bit & 31);
    pix_out[i] = pix;
    lab_out[i] = (int)labels[pix];
}

size_t sampler_ws_bytes(int64_t HW) { return SamplerWs(HW).total + 256; }

template <typename T>
static int run_sampler(const T* labels, int64_t HW, int n, const float* u, void* ws, int64_t* pix_out, int* lab_out,
                       cudaStream_t stream) {
    SamplerWs L(HW);
    char* w = static_cast<char*>(ws);
    uint32_t* words = reinterpret_cast<uint32_t*>(w + L.words);
    uint32_t* counts = reinterpret_cast<uint32_t*>(w + L.counts);
    label_words_kernel<T><<<L.nblocks, 256, 0, stream>>>(HW, labels, words, counts); note_launch();
    const size_t smem = (size_t)(L.nblocks + 1) * 4;
    if (smem > 200 * 1024) return ISR_ERR_UNSUPPORTED;  // > 200 M pixels
    auto kern = label_select_kernel<T>;
    if (smem > 40 * 1024) ISR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(n + 255) / 256, 256, smem, stream>>>(HW, n, u, words, counts, L.nblocks, labels, pix_out, lab_out); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_sampler(const void* labels, int label_bytes, int64_t HW, int n, const float* u, void* ws, int64_t* pix_out,
                   int* lab_out, cudaStream_t stream) {
    if (n <= 0 || HW <= 0) return ISR_OK;
    switch (label_bytes) {
        case 1: return run_sampler(static_cast<const int8_t*>(labels), HW, n, u, ws, pix_out, lab_out, stream);
        case 2: return run_sampler(static_cast<const int16_t*>(labels), HW, n, u, ws, pix_out, lab_out, stream);
        case 4: return run_sampler(static_cast<const int32_t*>(labels), HW, n, u, ws, pix_out, lab_out, stream);
        case 8: return run_sampler(static_cast<const int64_t*>(labels), HW, n, u, ws, pix_out, lab_out, stream);
        default: return ISR_ERR_INVALID_ARG;
    }
}

}  // namespace isr
