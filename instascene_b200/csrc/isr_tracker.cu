// isr_tracker.cu -- per-view Gaussian-tracker extraction on the device.
// Reference: spatial_track/modules/init_tracker.py:16-47 (get_segmap_gaussians): the (gaussian, pixel) pair list of
// one rendered view is moved to Python, and for every mask id of the view's segmentation map a Python set of the
// Gaussian ids whose pairs fall on that mask is built (`set(gaus_ids[valid_mask].tolist())`), masks with fewer than
// 50 distinct Gaussians are dropped; plus the set of all Gaussian ids of the frame.
//
// Here the pair list never leaves HBM.  "Set of Gaussian ids" = one bitmap row of P bits:
//   row 0      : every Gaussian that appears in the pair list            (frame_gaussian_ids)
//   row r >= 1 : Gaussians with a pair on a pixel whose dense mask row is r (mask_info[mask id of row r])
// mark  : one pass over the pairs, test-then-atomicOr (a Gaussian re-appears for many pixels: most pairs find their
//         bit already set and issue no atomic),
// count : popcount per row (the `len(set(...)) < 50` test runs on the host on K integers),
// fill  : per kept row an ordered compaction of the set bits -> ascending Gaussian ids (CSR), block-wide scan.
// Integer work, HBM/atomic bound: algorithmic bytes = 8 G (pairs) + 4 G (gathered mask row) + K*P/8 (bitmap, twice).
#include "isr_common.cuh"

namespace isr {

struct TrackerWs {
    size_t words_per_row, bitmap, total;
    TrackerWs(int P, int K) {
        words_per_row = ((size_t)(P > 0 ? P : 0) + 31) / 32;
        bitmap = 0;
        total = align_up((size_t)(K > 0 ? K : 0) * words_per_row * 4, 256);
    }
};

size_t tracker_ws_bytes(int P, int K) { return TrackerWs(P, K).total; }

__global__ void __launch_bounds__(256)
tracker_mark_kernel(const int2* __restrict__ pairs, int64_t n_pairs, const int* __restrict__ seg_rows, int64_t HW, int P,
                    int K, uint32_t* __restrict__ bitmap, size_t wpr) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
        const int2 pr = __ldg(pairs + i);
        const int g = pr.x;
        if (g < 0 || g >= P) continue;  // (the reference would index out of range; such pairs are never produced)
        const uint32_t bit = 1u << (g & 31);
        uint32_t* w0 = bitmap + (g >> 5);
        if (!(*w0 & bit)) atomicOr(w0, bit);
        const int64_t pix = pr.y;
        if (pix < 0 || pix >= HW) continue;
        const int r = __ldg(seg_rows + pix);
        if (r <= 0 || r >= K) continue;  // row 0 = background (mask id 0 is skipped, init_tracker.py:36-37)
        uint32_t* wr = bitmap + (size_t)r * wpr + (g >> 5);
        if (!(*wr & bit)) atomicOr(wr, bit);
    }
}

__global__ void __launch_bounds__(256)
tracker_count_kernel(const uint32_t* __restrict__ bitmap, size_t wpr, int* __restrict__ counts) {
    const int row = blockIdx.y;
    const uint32_t* rowp = bitmap + (size_t)row * wpr;
    int c = 0;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < wpr; w += (size_t)gridDim.x * blockDim.x)
        c += __popc(rowp[w]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(counts + row, c);
}

// One CTA per row; rows with row_offsets[row] < 0 are skipped.  Ordered compaction: every iteration takes blockDim
// consecutive words, an exclusive block scan of their popcounts gives each thread its output position.
constexpr int kFillThreads = 1024;
__global__ void __launch_bounds__(kFillThreads)
tracker_fill_kernel(const uint32_t* __restrict__ bitmap, size_t wpr, const int64_t* __restrict__ row_offsets,
                    int* __restrict__ out_ids) {
    __shared__ int warp_sums[kFillThreads / 32];
    __shared__ int chunk_total;
    const int row = blockIdx.x;
    const int64_t off = row_offsets[row];
    if (off < 0) return;
    const uint32_t* rowp = bitmap + (size_t)row * wpr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t base = off;
    for (size_t w0 = 0; w0 < wpr; w0 += kFillThreads) {
        const size_t w = w0 + threadIdx.x;
        uint32_t word = (w < wpr) ? rowp[w] : 0u;
        const int c = __popc(word);
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int s = warp_sums[lane];
            int si = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, si, d);
                if (lane >= d) si += t;
            }
            warp_sums[lane] = si - s;  // exclusive
            if (lane == 31) chunk_total = si;
        }
        __syncthreads();
        int64_t pos = base + warp_sums[warp] + (incl - c);
        const int gbase = (int)(w << 5);
        while (word) {
            const int b = __ffs(word) - 1;
            word &= word - 1;
            out_ids[pos++] = gbase + b;
        }
        base += chunk_total;
        __syncthreads();  // warp_sums / chunk_total are rewritten by the next iteration
    }
}

int launch_tracker_mark(const int* pairs, int64_t n_pairs, const int* seg_rows, int64_t HW, int P, int K, void* ws,
                        int* counts, cudaStream_t stream) {
    TrackerWs L(P, K);
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + L.bitmap);
    ISR_CUDA_TRY(cudaMemsetAsync(bitmap, 0, (size_t)K * L.words_per_row * 4, stream));
    ISR_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)K * sizeof(int), stream));
    if (n_pairs > 0) {
        const int64_t want = (n_pairs + 255) / 256;
        const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);  // 16 CTAs of 256 threads per SM, grid-stride
        tracker_mark_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const int2*>(pairs), n_pairs, seg_rows, HW, P, K,
                                                        bitmap, L.words_per_row); note_launch();
    }
    if (L.words_per_row > 0) {
        const size_t per_block = 256 * 8;
        int bx = (int)((L.words_per_row + per_block - 1) / per_block);
        if (bx < 1) bx = 1;
        tracker_count_kernel<<<dim3(bx, K), 256, 0, stream>>>(bitmap, L.words_per_row, counts); note_launch();
    }
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_tracker_fill(int P, int K, const void* ws, const int64_t* row_offsets, int* out_ids, cudaStream_t stream) {
    TrackerWs L(P, K);
    const uint32_t* bitmap = reinterpret_cast<const uint32_t*>(static_cast<const char*>(ws) + L.bitmap);
    if (L.words_per_row == 0) return ISR_OK;
    tracker_fill_kernel<<<K, kFillThreads, 0, stream>>>(bitmap, L.words_per_row, row_offsets, out_ids); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

}  // namespace isr
