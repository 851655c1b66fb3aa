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

bool entries_packed(int P) {
    static const bool plain = [] { const char* e = getenv("ISR_PLAIN_ENTRIES"); return e && e[0] == '1'; }();
    return !plain && P < (1 << kIdBits);
}

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

// One list entry: Gaussian id, plus (packed entries) the per-block footprint bits of tile (tile_x, tile_y): the test
// of isr::rect_may_touch / block_outside for the tile's 2 x 4 blocks of 8x4 pixels, with the conic evaluated
// incrementally over the block-centre grid (Q = a x^2 + (b y + d) x + (c y^2 + e y + f) per row) and UNCLIPPED block
// extents at the image border (a larger block only makes the test more conservative).
__device__ __forceinline__ uint32_t make_entry(uint32_t g, int tile_x, int tile_y, int packed, const float4 cr,
                                               const float4 q0, const float4 q1, float r2) {
    if (!packed) return g;
    const float X0 = (float)(tile_x * TILE), Y0 = (float)(tile_y * TILE);
    const float x0 = X0 + 3.5f - q1.z, y0 = Y0 + 1.5f - q1.w;  // centre of block (0,0) relative to the Gaussian's centre
    const float a2 = q0.x + q0.x, c2 = q0.z + q0.z;
    uint32_t bits = 0;
#pragma unroll
    for (int by = 0; by < 4; by++) {
        const float y = y0 + 4.0f * (float)by;
        const float py0 = Y0 + 4.0f * (float)by;
        const bool yin = !(cr.w < py0 || cr.y > py0 + 3.0f);
        const float lin = fmaf(q0.y, y, q0.w);                 // b y + d
        const float cst = fmaf(fmaf(q0.z, y, q1.x), y, q1.y);  // c y^2 + e y + f
        const float gy0 = fmaf(c2, y, q1.x);                   // 2 c y + e
        const float ey = fmaxf(fabsf(y) - 1.5f, 0.0f), ey2 = ey * ey;
#pragma unroll
        for (int bx = 0; bx < 2; bx++) {
            const float x = x0 + 8.0f * (float)bx;
            const float px0 = X0 + 8.0f * (float)bx;
            const bool xin = !(cr.z < px0 || cr.x > px0 + 7.0f);
            const float Qc = fmaf(fmaf(q0.x, x, lin), x, cst);
            const float gxx = fmaf(a2, x, lin), gyy = fmaf(q0.y, x, gy0);
            const bool out3d = Qc - fmaf(fabsf(gxx), 3.5f, fabsf(gyy) * 1.5f) > 0.02f;
            const float ex = fmaxf(fabsf(x) - 3.5f, 0.0f);
            const bool out2d = fmaf(ex, ex, ey2) > r2;
            if (xin && yin && !(out3d && out2d)) bits |= 1u << (by * 2 + bx);
        }
    }
    return g | (bits << kIdBits);
}

// Instance emission (DSR duplicateWithKeys, rasterizer_impl.cu:70-111) in depth order.  A warp owns 32 consecutive
// Gaussians of the depth order.  Footprints of up to 64 tiles (K1 mask): the lanes take consecutive instances of the
// warp's Gaussians -- each finds the owning Gaussian by a 5-step search over the warp's running counts (shuffles), the
// tile as the k-th set bit of that Gaussian's mask, evaluates the 8 per-block footprint bits of that tile and writes
// (tile id, Gaussian id | block bits << 24): coalesced stores, no per-Gaussian serial loop, all lanes busy.  Larger
// footprints are only queued here (id, offset) and expanded by emit_big_kernel with a whole CTA each, so that one
// screen-filling splat does not serialise thousands of instances on a single warp.  Tile t of a rectangle is numbered
// row-major (the reference's order); the order of one Gaussian's instances is irrelevant for the result (distinct
// tiles), the order ACROSS Gaussians is the depth order.
__global__ void __launch_bounds__(256)
emit_instances_kernel(int P, const uint32_t* __restrict__ order, const uint32_t* __restrict__ offsets,
                      const int* __restrict__ radii, const Splat* __restrict__ splats,
                      const unsigned long long* __restrict__ tile_mask, const uint32_t* __restrict__ tile_count,
                      const float4* __restrict__ cull4, const float4* __restrict__ cullq, int gx, int gy, int W, int H,
                      int packed, uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ gauss_ids,
                      uint32_t* __restrict__ big_count, uint2* __restrict__ big_list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t g = 0, off = 0, cnt = 0, rect = 0, w = 1;
    unsigned long long mask = 0ull;
    if (i < P) {
        g = order[i];
        cnt = tile_count[g];
        if (cnt > 0) {
            off = offsets[i];
            int mnx, mny, mxx, mxy;
            get_rect(splats[g].mx, splats[g].my, radii[g], gx, gy, mnx, mny, mxx, mxy);
            w = (uint32_t)(mxx - mnx);
            if ((mxx - mnx) * (mxy - mny) > 64) {  // expanded by emit_big_kernel
                big_list[atomicAdd(big_count, 1u)] = make_uint2(g, off);
                cnt = 0;
            } else {
                rect = (uint32_t)mnx | ((uint32_t)mny << 16);
                mask = tile_mask[g];
            }
        }
    }
    // running instance count over the warp's Gaussians (inclusive scan), c_excl = instances before this lane's
    uint32_t c_incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, c_incl, d);
        if (lane >= d) c_incl += v;
    }
    const uint32_t c_excl = c_incl - cnt;
    const uint32_t total = __shfl_sync(0xffffffffu, c_incl, 31);
    const uint32_t mlo = (uint32_t)mask, mhi = (uint32_t)(mask >> 32);
    for (uint32_t jb = 0; jb < total; jb += 32) {
        const uint32_t j = jb + lane;
        int lo = 0;  // owner = last lane whose exclusive count is <= j (lanes with no instances share their successor's)
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const uint32_t o = __shfl_sync(0xffffffffu, c_excl, lo + step);
            if (o <= j) lo += step;
        }
        const uint32_t g_o = __shfl_sync(0xffffffffu, g, lo), off_o = __shfl_sync(0xffffffffu, off, lo);
        const uint32_t ce_o = __shfl_sync(0xffffffffu, c_excl, lo);
        const uint32_t rect_o = __shfl_sync(0xffffffffu, rect, lo), w_o = __shfl_sync(0xffffffffu, w, lo);
        const uint32_t mlo_o = __shfl_sync(0xffffffffu, mlo, lo), mhi_o = __shfl_sync(0xffffffffu, mhi, lo);
        if (j >= total) continue;
        const int k = (int)(j - ce_o);
        unsigned long long m = (unsigned long long)mlo_o | ((unsigned long long)mhi_o << 32);
        for (int q = 0; q < k; q++) m &= m - 1;  // drop the k lowest set bits
        const int t = __ffsll((long long)m) - 1;
        // t / w without an integer division: (t + 0.5) / w is at least 0.5/w away from an integer
        const int ty = __float2int_rd(((float)t + 0.5f) * (1.0f / (float)w_o)), tx = t - ty * (int)w_o;
        const int tile_x = (int)(rect_o & 0xffffu) + tx, tile_y = (int)(rect_o >> 16) + ty;
        float4 cr = make_float4(0.f, 0.f, 0.f, 0.f), q0 = cr, q1 = cr;
        float r2 = 0.0f;
        if (packed) {
            cr = __ldg(cull4 + g_o);
            const float4* q = cullq + (size_t)g_o * 3;
            q0 = __ldg(q); q1 = __ldg(q + 1); r2 = __ldg(q + 2).x;
        }
        tile_keys[off_o + k] = (uint32_t)(tile_y * gx + tile_x);
        gauss_ids[off_o + k] = make_entry(g_o, tile_x, tile_y, packed, cr, q0, q1, r2);
    }
}

// Footprints of more than 64 tiles: one CTA per queued Gaussian, every tile of its getRect rectangle (row-major).
__global__ void __launch_bounds__(256)
emit_big_kernel(const uint32_t* __restrict__ big_count, const uint2* __restrict__ big_list,
                const int* __restrict__ radii, const Splat* __restrict__ splats, const float4* __restrict__ cull4,
                const float4* __restrict__ cullq, int gx, int gy, int W, int H, int packed,
                uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ gauss_ids) {
    const uint32_t n_big = *big_count;
    for (uint32_t e = blockIdx.x; e < n_big; e += gridDim.x) {
        const uint2 it = big_list[e];
        const uint32_t g = it.x, off = it.y;
        int mnx, mny, mxx, mxy;
        get_rect(splats[g].mx, splats[g].my, radii[g], gx, gy, mnx, mny, mxx, mxy);
        const int w = mxx - mnx, n = w * (mxy - mny);
        const float inv_w = 1.0f / (float)w;
        float4 cr = make_float4(0.f, 0.f, 0.f, 0.f), q0 = cr, q1 = cr;
        float r2 = 0.0f;
        if (packed) {
            cr = __ldg(cull4 + g);
            const float4* q = cullq + (size_t)g * 3;
            q0 = __ldg(q); q1 = __ldg(q + 1); r2 = __ldg(q + 2).x;
        }
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            const int ty = __float2int_rd(((float)t + 0.5f) * inv_w), tx = t - ty * w;
            tile_keys[off + t] = (uint32_t)((mny + ty) * gx + (mnx + tx));
            gauss_ids[off + t] = make_entry(g, mnx + tx, mny + ty, packed, cr, q0, q1, r2);
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
    uint32_t* big_count = reinterpret_cast<uint32_t*>(g + gl.counters) + 1;
    uint2* big_list = reinterpret_cast<uint2*>(g + gl.big_list);
    const int packed = entries_packed(P) ? 1 : 0;
    ISR_CUDA_TRY(cudaMemsetAsync(big_count, 0, sizeof(uint32_t), stream));
    emit_instances_kernel<<<(P + 255) / 256, 256, 0, stream>>>(
        P, reinterpret_cast<const uint32_t*>(g + gl.order), reinterpret_cast<const uint32_t*>(g + gl.offsets), a.radii,
        reinterpret_cast<const Splat*>(g + gl.splat), reinterpret_cast<const unsigned long long*>(g + gl.tmask),
        reinterpret_cast<const uint32_t*>(g + gl.tcount), reinterpret_cast<const float4*>(g + gl.cull),
        reinterpret_cast<const float4*>(g + gl.cullq), gx, gy, a.W, a.H, packed, tile_keys_alt, point_list_alt, big_count,
        big_list);
    ISR_CUDA_TRY(cudaGetLastError());
    emit_big_kernel<<<592, 256, 0, stream>>>(big_count, big_list, a.radii, reinterpret_cast<const Splat*>(g + gl.splat),
                                             reinterpret_cast<const float4*>(g + gl.cull),
                                             reinterpret_cast<const float4*>(g + gl.cullq), gx, gy, a.W, a.H, packed,
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
