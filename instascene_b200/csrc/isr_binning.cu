// isr_binning.cu -- tile binning and per-tile depth ordering.
//
// Reference (DSR/cuda_rasterizer/rasterizer_impl.cu:283-324): scan tiles_touched in Gaussian-id order, emit
// one 64-bit key (tile<<32 | depth bits) per (Gaussian, tile) instance, ONE stable radix sort of R 12-byte
// pairs over 32+log2(tiles) bits (6 onesweep passes at 1080p), then identifyTileRanges.
//
// Here:
//   0. (K1) per Gaussian, a 64-bit mask of the tiles of its getRect rectangle that its footprint can reach at all
//      (conservative conic / low-pass-disk test, isr::rect_may_touch): the reference's square of side 2*radius
//      emits many tiles an elongated or small splat provably never touches
//   1. stable sort of the P Gaussians by depth bits (culled ones carry key 0xFFFFFFFF and sink to the end)
//   2. exclusive scan of the emitted-tile counts in that depth order -> instance offsets, R_emit
//   3. emission of (tile id, Gaussian id) in depth order for the tiles in the mask
//   4. stable sort of the instances by tile id ONLY (13 bits at 1080p = 2 onesweep passes of 8-byte pairs)
//   5. tile ranges from the sorted tile ids
// Because both sorts are stable, instances of a tile end up ordered by (depth bits, Gaussian id): each tile list
// is exactly the reference's list (ties: SURVEY.md Q2) minus entries that contribute to none of the tile's pixels
// -- skipping those never changes a result -- while the instance-sized traffic drops by an order of magnitude.
// The reference's num_rendered (sum of tiles_touched) is still computed and reported at the boundary.
// The radix-sort/scan primitives are CUB (CUDA toolkit), as in the reference.
#include <cub/cub.cuh>

#include "isr_common.cuh"

namespace isr {

struct TilesInDepthOrder {
    const uint32_t* tiles;
    const uint32_t* order;
    __host__ __device__ uint32_t operator()(int i) const { return tiles[order[i]]; }
};

size_t sort_temp_bytes_gauss(int P) {
    if (P <= 0) return 256;
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, P, 0, 32);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, P + 1);
    size_t c = 0;
    cub::DeviceReduce::Sum(nullptr, c, (const uint32_t*)nullptr, (uint32_t*)nullptr, P);
    a = a > b ? a : b;
    return align_up((a > c ? a : c) + 256, 256);
}

size_t sort_temp_bytes_inst(int64_t R, int num_tiles) {
    if (R <= 0) return 256;
    size_t a = 0;
    int bits = 1;
    while ((1 << bits) < num_tiles) bits++;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)R, 0, bits);
    return align_up(a + 256, 256);
}

__global__ void iota_kernel(int n, uint32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}

// gathered[i] = tiles[order[i]] for i < P, gathered[P] = 0  (so that an exclusive scan over P+1 items leaves
// R in offsets[P])
__global__ void gather_tiles_kernel(int P, const uint32_t* __restrict__ tiles, const uint32_t* __restrict__ order,
                                    uint32_t* __restrict__ gathered) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) gathered[i] = tiles[order[i]];
    else if (i == P) gathered[i] = 0;
}

// out[0] = the reference's num_rendered (all tiles of every getRect rectangle), out[1] = emitted instances
__global__ void copy_count_kernel(const uint32_t* __restrict__ offsets, int P, const uint32_t* __restrict__ tile_total,
                                  int64_t* __restrict__ out) {
    out[0] = (int64_t)*tile_total;
    out[1] = (int64_t)offsets[P];
}

// Instance emission (DSR duplicateWithKeys, rasterizer_impl.cu:70-111) in depth order.  A warp takes 32 consecutive
// Gaussians; each lane fetches one Gaussian's rectangle, tile mask and offset, then the warp writes the Gaussians'
// (tile id, Gaussian id) runs one after the other with all lanes (contiguous, fully used sectors) instead of 32
// lanes each walking its own run with 4-byte scattered stores.  Tile t of the rectangle (row-major, the reference's
// order) is emitted iff bit t of the mask is set; rectangles of more than 64 tiles are emitted whole.
__global__ void __launch_bounds__(256)
emit_instances_kernel(int P, const uint32_t* __restrict__ order, const uint32_t* __restrict__ offsets,
                      const int* __restrict__ radii, const Splat* __restrict__ splats,
                      const unsigned long long* __restrict__ tile_mask, int gx, int gy,
                      uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ gauss_ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t g = 0, off = 0;
    int mnx = 0, mny = 0, w = 0, n = 0;
    unsigned long long mask = 0ull;
    float inv_w = 0.0f;
    if (i < P) {
        g = order[i];
        const int r = radii[g];
        if (r > 0) {
            int mxx, mxy;
            get_rect(splats[g].mx, splats[g].my, r, gx, gy, mnx, mny, mxx, mxy);
            w = mxx - mnx;
            n = w * (mxy - mny);
            inv_w = 1.0f / (float)w;
            off = offsets[i];
            mask = tile_mask[g];
            if (mask == 0ull) n = 0;
        }
    }
    unsigned todo = __ballot_sync(0xffffffffu, n > 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t g_s = __shfl_sync(0xffffffffu, g, src), off_s = __shfl_sync(0xffffffffu, off, src);
        const int mnx_s = __shfl_sync(0xffffffffu, mnx, src), mny_s = __shfl_sync(0xffffffffu, mny, src);
        const int w_s = __shfl_sync(0xffffffffu, w, src), n_s = __shfl_sync(0xffffffffu, n, src);
        const float iw_s = __shfl_sync(0xffffffffu, inv_w, src);
        const unsigned long long m_s = __shfl_sync(0xffffffffu, mask, src);
        if (n_s <= 64) {
            for (int t = lane; t < n_s; t += 32) {
                if ((m_s >> t) & 1ull) {
                    const int ty = __float2int_rd(((float)t + 0.5f) * iw_s), tx = t - ty * w_s;
                    const uint32_t dst = off_s + (uint32_t)__popcll(m_s & ((1ull << t) - 1ull));
                    tile_keys[dst] = (uint32_t)((mny_s + ty) * gx + (mnx_s + tx));
                    gauss_ids[dst] = g_s;
                }
            }
        } else {
            for (int t = lane; t < n_s; t += 32) {  // t-th tile of the rectangle, row-major
                // t / w without an integer division: (t + 0.5) / w is at least 0.5/w away from an integer, far more
                // than the fp32 error for any t below ~1e6 tiles
                const int ty = __float2int_rd(((float)t + 0.5f) * iw_s), tx = t - ty * w_s;
                tile_keys[off_s + t] = (uint32_t)((mny_s + ty) * gx + (mnx_s + tx));
                gauss_ids[off_s + t] = g_s;
            }
        }
    }
}

// DSR identifyTileRanges (rasterizer_impl.cu:116-138) on 32-bit tile keys.
__global__ void tile_ranges_kernel(int64_t R, const uint32_t* __restrict__ keys, uint2* __restrict__ ranges) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t cur = keys[i];
    if (i == 0) ranges[cur].x = 0;
    else {
        const uint32_t prev = keys[i - 1];
        if (cur != prev) {
            ranges[prev].y = (uint32_t)i;
            ranges[cur].x = (uint32_t)i;
        }
    }
    if (i == R - 1) ranges[cur].y = (uint32_t)R;
}

// Phase A tail: depth order + offsets + R -> pinned host.
int launch_depth_order_and_offsets(const IsrForwardArgs& a, cudaStream_t stream) {
    const int P = a.P;
    GeomLayout gl(P);
    char* g = static_cast<char*>(a.geom);
    uint32_t* order = reinterpret_cast<uint32_t*>(g + gl.order);
    uint32_t* order_alt = reinterpret_cast<uint32_t*>(g + gl.order_alt);
    uint32_t* keys = reinterpret_cast<uint32_t*>(g + gl.depth_key);
    uint32_t* keys_alt = reinterpret_cast<uint32_t*>(g + gl.keys_alt);
    uint32_t* tiles = reinterpret_cast<uint32_t*>(g + gl.tiles);
    uint32_t* tcount = reinterpret_cast<uint32_t*>(g + gl.tcount);
    uint32_t* tile_total = reinterpret_cast<uint32_t*>(g + gl.counters);
    uint32_t* offsets = reinterpret_cast<uint32_t*>(g + gl.offsets);
    void* temp = g + gl.sort_temp;
    size_t temp_bytes = gl.sort_temp_bytes;
    if (a.num_rendered_host != nullptr)  // only the boundary's num_rendered needs the reference's count
        ISR_CUDA_TRY(cub::DeviceReduce::Sum(temp, temp_bytes, tiles, tile_total, P, stream));
    temp_bytes = gl.sort_temp_bytes;
    iota_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, order_alt);
    ISR_CUDA_TRY(cudaGetLastError());
    // stable: equal depth bits keep ascending Gaussian id
    ISR_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_alt, order_alt, order, P, 0, 32, stream));
    // keys_alt now holds sorted keys (unused afterwards) -> reuse it for the gathered tile counts
    gather_tiles_kernel<<<(P + 1 + 255) / 256, 256, 0, stream>>>(P, tcount, order, keys_alt);
    ISR_CUDA_TRY(cudaGetLastError());
    temp_bytes = gl.sort_temp_bytes;
    ISR_CUDA_TRY(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, keys_alt, offsets, P + 1, stream));
    if (a.num_rendered_host != nullptr) {
        // both counts -> int64[2] on the host; widened on the device into the scratch first
        int64_t* wide = reinterpret_cast<int64_t*>(temp);
        copy_count_kernel<<<1, 1, 0, stream>>>(offsets, P, tile_total, wide);
        ISR_CUDA_TRY(cudaGetLastError());
        ISR_CUDA_TRY(cudaMemcpyAsync(a.num_rendered_host, wide, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    }
    return ISR_OK;
}

// Phase B head: emission, tile sort, ranges.
int launch_binning(const IsrForwardArgs& a, int64_t R, cudaStream_t stream) {
    const int P = a.P;
    const int gx = (a.W + TILE - 1) / TILE, gy = (a.H + TILE - 1) / TILE;
    const int num_tiles = gx * gy;
    GeomLayout gl(P);
    ImageLayout il(a.W, a.H);
    BinLayout bl(P, R, a.W, a.H);
    char* g = static_cast<char*>(a.geom);
    char* im = static_cast<char*>(a.image);
    char* b = static_cast<char*>(a.binning);
    uint2* ranges = reinterpret_cast<uint2*>(im + il.ranges);
    ISR_CUDA_TRY(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)num_tiles, stream));
    if (R <= 0 || P <= 0) return ISR_OK;
    uint32_t* point_list = reinterpret_cast<uint32_t*>(b + bl.point_list);
    uint32_t* point_list_alt = reinterpret_cast<uint32_t*>(b + bl.point_list_alt);
    uint32_t* tile_keys = reinterpret_cast<uint32_t*>(b + bl.tile_keys);
    uint32_t* tile_keys_alt = reinterpret_cast<uint32_t*>(b + bl.tile_keys_alt);
    emit_instances_kernel<<<(P + 255) / 256, 256, 0, stream>>>(
        P, reinterpret_cast<const uint32_t*>(g + gl.order), reinterpret_cast<const uint32_t*>(g + gl.offsets), a.radii,
        reinterpret_cast<const Splat*>(g + gl.splat), reinterpret_cast<const unsigned long long*>(g + gl.tmask), gx, gy,
        tile_keys_alt, point_list_alt);
    ISR_CUDA_TRY(cudaGetLastError());
    int bits = 1;
    while ((1 << bits) < num_tiles) bits++;
    size_t temp_bytes = bl.temp_bytes;
    ISR_CUDA_TRY(cub::DeviceRadixSort::SortPairs(b + bl.temp, temp_bytes, tile_keys_alt, tile_keys, point_list_alt,
                                                 point_list, (int)R, 0, bits, stream));
    tile_ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(R, tile_keys, ranges);
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

}  // namespace isr
