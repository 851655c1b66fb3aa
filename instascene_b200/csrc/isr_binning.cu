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
//   3. STABLE COUNTING PARTITION of the instances by tile, fused with their emission -- no (key, value) arrays, no
//      sort passes over the instances, nothing sized by R on the host:
//        a. bin_count_kernel: the depth order is cut into `chunks` runs of ~R/chunks instances; one CTA per run
//           enumerates its instances and counts them per tile in shared memory (one 32-bit counter per tile) -> a
//           [chunks][tiles] table
//        b. bin_colscan_kernel / bin_tilebase_kernel / bin_addbase_kernel: exclusive scan down every tile column and
//           across the tile totals: table[c][t] becomes the list position of chunk c's first instance in tile t, the
//           totals become the tile ranges (identifyTileRanges for free)
//        c. bin_scatter_kernel: one CTA per run re-enumerates it in depth order into a shared-memory ring, 512
//           instances per round, and writes every list entry (Gaussian id + the 8 per-block footprint bits) straight
//           to its final position: start of (run, tile) + shared-memory cursor of the tile + instances of the same
//           round in earlier warps (one count byte per warp in a shared-memory word per tile) + rank in the warp
//           (ISR_BIN_UNORDERED=1: unordered slots + a sort of every (run, tile) segment by depth rank instead; same
//           lists, measured equal)
//      Instances of a tile end up ordered by (depth bits, Gaussian id): each tile list is exactly the reference's list
//      (ties: SURVEY.md Q2) minus entries that contribute to none of the tile's pixels -- skipping those never
//      changes a result.  The kernels read R from device memory, so phase B has no host-side dependence on it beyond
//      the capacity of the list buffer.
//   (fallback when the per-tile tables do not fit in shared memory, > ~34k tiles (beyond 4K resolution): emission of (tile id, entry) pairs
//    + CUB radix sort by tile id + tile ranges, as in round 1)
// The reference's num_rendered (sum of tiles_touched) is still computed and reported at the boundary.
// Phase A's depth sort and offset scan use CUB (CUDA toolkit), as the reference does for its single big sort.
#include <cub/cub.cuh>

#include "isr_common.cuh"

namespace isr {

bool entries_packed(int P) {
    static const bool plain = [] { const char* e = getenv("ISR_PLAIN_ENTRIES"); return e && e[0] == '1'; }();
    return !plain && P < (1 << kIdBits);
}

size_t scatter_smem_bytes(int num_tiles);
size_t scatter_u_smem_bytes(int num_tiles);

static int bin_sm_count() {
    static const int sm_count = [] {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;  // B200
        return n;
    }();
    return sm_count;
}

BinChunks bin_chunks(int num_tiles, int64_t R) {
    const int sm_count = bin_sm_count();
    // ISR_BIN_UNORDERED=1: unordered scatter + sort of the (run, tile) segments instead of the in-order ranking kernel
    // (measured equal overall: 243 + 111 us against 366 us, two more kernels and 2R words of workspace; off by default)
    static const bool ordered = [] { const char* e = getenv("ISR_BIN_UNORDERED"); return !(e && e[0] == '1'); }();
    static const bool force_sort = [] { const char* e = getenv("ISR_BIN_SORT"); return e && e[0] == '1'; }();  // test hook
    BinChunks bc;
    const int nt = num_tiles > 0 ? num_tiles : 1;
    bc.smem_count = align_up((size_t)nt * 4, 16);
    bc.unordered = !ordered;
    bc.smem_scatter = bc.unordered ? scatter_u_smem_bytes(nt) : scatter_smem_bytes(nt);
    // per SM: 227 KB opt-in, 1 KB of system reservation per resident CTA; one chunk = one CTA of bin_scatter_kernel
    const size_t budget = 226 * 1024;
    int per_sm = (int)(budget / (bc.smem_scatter + 1024 + 1280 /* static */));
    if (per_sm > 5) per_sm = 5;
    bc.chunks = (force_sort || per_sm < 1 || num_tiles > 65535) ? 0 : sm_count * per_sm;
    // the 16-bit per-tile cursors of bin_scatter_kernel count at most one instance per Gaussian of the chunk: a chunk
    // must start fewer than 65536 instances
    if (bc.chunks > 0 && R > (int64_t)bc.chunks * 60000) {
        const int64_t mult = (R + (int64_t)bc.chunks * 60000 - 1) / ((int64_t)bc.chunks * 60000);
        if (mult > 64) bc.chunks = 0;
        else bc.chunks *= (int)mult;
    }
    return bc;
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
    a = a > b ? a : b;
    return align_up(a + 256, 256);
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

// gathered[i] = tcount[order[i]] for i < P, gathered[P] = 0  (so that an exclusive scan over P+1 items leaves
// R in offsets[P]).  Also accumulates, in 64 bits, the reference's num_rendered (sum of tiles_touched) and the
// emitted-instance total (the 32-bit scan would wrap silently beyond 2^32 instances), and writes what the counting
// partition needs per Gaussian as ONE record in depth order (so that the per-chunk warps stream it sequentially instead
// of chasing order[i] -> per-Gaussian arrays): id, emitted tile count (bit 31: footprint of more than 64 tiles = every
// tile of the rectangle), getRect origin, rectangle width, and the K1 footprint mask.
__global__ void gather_tiles_kernel(int P, const uint32_t* __restrict__ order,
                                    const uint32_t* __restrict__ tiles_touched, const uint4* __restrict__ tile_foot,
                                    uint32_t* __restrict__ gathered, uint4* __restrict__ bin_rec,
                                    unsigned long long* __restrict__ bin_mask, uint32_t* __restrict__ rank,
                                    unsigned long long* __restrict__ totals /*[0] sum tiles_touched, [1] sum tcount*/) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long a = 0ull, b = 0ull;
    if (i < P) {
        const uint32_t g = order[i];
        const uint4 f0 = __ldg(tile_foot + 2 * (size_t)g), f1 = __ldg(tile_foot + 2 * (size_t)g + 1);  // one 32-byte sector
        const uint32_t c = f0.x & 0x7fffffffu;
        gathered[i] = c;
        b = c;
        a = tiles_touched[i];
        bin_rec[i] = make_uint4(g, f0.x, f0.y, c ? f0.z : 1u);
        bin_mask[i] = (unsigned long long)f1.x | ((unsigned long long)f1.y << 32);
        rank[g] = (uint32_t)i;
    } else if (i == P) {
        gathered[i] = 0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    // one pair of atomics per CTA (per warp they were 125 k same-address atomics: the bottleneck of this kernel)
    __shared__ unsigned long long s_tot[2][8];
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_tot[0][wid] = a; s_tot[1][wid] = b; }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned long long t = 0ull;
#pragma unroll
        for (int w = 0; w < 8; w++) t += s_tot[threadIdx.x][w];
        if (t) atomicAdd(totals + threadIdx.x, t);
    }
}

// out[0] = the reference's num_rendered (all tiles of every getRect rectangle), out[1] = emitted instances (64-bit sums)
__global__ void copy_count_kernel(const unsigned long long* __restrict__ totals, int64_t* __restrict__ out) {
    out[0] = (int64_t)totals[0];
    out[1] = (int64_t)totals[1];
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
                      const uint4* __restrict__ tile_foot, const uint32_t* __restrict__ tile_count,
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
                const uint4 f1 = tile_foot[2 * (size_t)g + 1];
                mask = (unsigned long long)f1.x | ((unsigned long long)f1.y << 32);
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
    unsigned long long* totals = reinterpret_cast<unsigned long long*>(g + gl.counters);  // [0], [1]; see GeomLayout
    uint32_t* offsets = reinterpret_cast<uint32_t*>(g + gl.offsets);
    void* temp = g + gl.sort_temp;
    size_t temp_bytes = gl.sort_temp_bytes;
    ISR_CUDA_TRY(cudaMemsetAsync(totals, 0, 32, stream));  // both 64-bit totals + big_list count + overflow flag
    iota_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, order_alt); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    // stable: equal depth bits keep ascending Gaussian id
    ISR_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_alt, order_alt, order, P, 0, 32, stream));
    // keys_alt now holds sorted keys (unused afterwards) -> reuse it for the gathered tile counts
    gather_tiles_kernel<<<(P + 1 + 255) / 256, 256, 0, stream>>>(
        P, order, tiles, reinterpret_cast<const uint4*>(g + gl.tfoot), keys_alt, reinterpret_cast<uint4*>(g + gl.bin_rec),
        reinterpret_cast<unsigned long long*>(g + gl.bin_mask), reinterpret_cast<uint32_t*>(g + gl.rank), totals); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    temp_bytes = gl.sort_temp_bytes;
    ISR_CUDA_TRY(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, keys_alt, offsets, P + 1, stream));
    if (a.num_rendered_host != nullptr) {
        // both counts -> int64[2] on the host
        int64_t* wide = reinterpret_cast<int64_t*>(temp);
        copy_count_kernel<<<1, 1, 0, stream>>>(totals, wide); note_launch();
        ISR_CUDA_TRY(cudaGetLastError());
        ISR_CUDA_TRY(cudaMemcpyAsync(a.num_rendered_host, wide, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    }
    return ISR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Stable counting partition by tile, fused with emission (steps 3a-3c of the header).
// ---------------------------------------------------------------------------------------------------------
// First index i in [0, n] with a[i] >= target (a non-decreasing, a[n] readable); all lanes cooperate: 32-ary search.
__device__ __forceinline__ int warp_lower_bound(const uint32_t* __restrict__ a, int n, uint32_t target, int lane) {
    int lo = 0, hi = n;  // the answer lies in [lo, hi]
    while (hi - lo > 32) {
        const int step = (hi - lo + 32) / 33;
        const int p = lo + (lane + 1) * step - 1;
        const bool ge = p < hi ? (__ldg(a + p) >= target) : true;
        const unsigned b = __ballot_sync(0xffffffffu, ge);
        if (b == 0u) {
            lo += 32 * step;
        } else {
            const int first = __ffs(b) - 1;
            const int p_first = lo + (first + 1) * step - 1;
            if (p_first < hi) hi = p_first;
            lo += first * step;
        }
    }
    const int p = lo + lane;
    const bool ge = p < hi ? (__ldg(a + p) >= target) : true;
    return lo + __ffs(__ballot_sync(0xffffffffu, ge)) - 1;
}

struct BinArgs {
    int P, gx, gy, W, H, num_tiles, chunks, packed;
    const uint32_t* offsets;         // [P+1] exclusive scan of the emitted tile counts in depth order; offsets[P] = R
    const uint4* rec;                // [P] depth order (gather_tiles_kernel)
    const unsigned long long* mask;  // [P] depth order
    const float4* cull4;
    const float4* cullq;
};

// Depth-order positions [i0, i1) of chunk c: the Gaussians whose first instance lies in [c*q, (c+1)*q), q = ceil(R/chunks).
__device__ __forceinline__ void chunk_range(const BinArgs& b, int c, int lane, int& i0, int& i1) {
    const uint32_t R = __ldg(b.offsets + b.P);
    const uint32_t q = max(1u, (R + (uint32_t)b.chunks - 1u) / (uint32_t)b.chunks);
    const unsigned long long lo = (unsigned long long)c * q, hi = lo + q;
    // Everything from the first position whose start offset equals R on has no instances (that tail holds every culled
    // Gaussian: hundreds of thousands of positions that must not land on the last chunk's single warp).
    if (lo >= R) {
        i0 = i1 = b.P;
        return;
    }
    i0 = warp_lower_bound(b.offsets, b.P, (uint32_t)lo, lane);
    i1 = warp_lower_bound(b.offsets, b.P, (uint32_t)(hi < R ? hi : R), lane);
}

// One batch = 32 consecutive Gaussians of the depth order, lane l holding Gaussian l.  bin_count_kernel walks a chunk
// with a CTA of kBinWarps warps, warp w taking batches w, w + kBinWarps, ... (counting needs no order).
constexpr int kBinWarps = 8;

struct BinBatch {
    uint32_t g, cnt, big, w;   // Gaussian id, emitted tiles (0: none / past the end), footprint > 64 tiles, rectangle width
    int mnx, mny;              // getRect origin
    unsigned long long mask;   // K1 footprint mask (small footprints)
};

__device__ __forceinline__ BinBatch load_batch(const BinArgs& b, int ib, int i1, int lane) {
    BinBatch bb;
    uint4 rec = make_uint4(0u, 0u, 0u, 1u);
    bb.mask = 0ull;
    const bool has_g = ib + lane < i1;
    if (has_g) { rec = __ldg(b.rec + ib + lane); bb.mask = __ldg(b.mask + ib + lane); }
    bb.g = rec.x;
    bb.cnt = has_g ? (rec.y & 0x7fffffffu) : 0u;
    bb.big = bb.cnt ? (rec.y >> 31) : 0u;
    bb.mnx = (int)(rec.z & 0xffffu);
    bb.mny = (int)(rec.z >> 16);
    bb.w = rec.w;
    return bb;
}

// tile (row-major index t of a rectangle of width w at (mnx, mny)) -> global tile id; t / w without an integer
// division: (t + 0.5) / w is at least 0.5/w away from an integer
__device__ __forceinline__ uint32_t rect_tile(int t, float inv_w, int w, int mnx, int mny, int gx) {
    const int ty = __float2int_rd(((float)t + 0.5f) * inv_w), tx = t - ty * w;
    return (uint32_t)((mny + ty) * gx + mnx + tx);
}

// 3a: per-(chunk, tile) instance counts.  One CTA per chunk; cnt[] = one 32-bit counter per tile in shared memory,
// shared by the CTA's warps (counting needs no order).  Every lane walks the set bits of its own Gaussian's footprint
// mask; the rare footprints of more than 64 tiles (every tile of the rectangle) are walked by the whole warp.
__global__ void __launch_bounds__(32 * kBinWarps) bin_count_kernel(const BinArgs b, uint32_t* __restrict__ table) {
    extern __shared__ uint32_t cnt[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = blockIdx.x;
    for (int t = threadIdx.x; t < b.num_tiles; t += blockDim.x) cnt[t] = 0u;
    __syncthreads();
    int i0, i1;
    chunk_range(b, c, lane, i0, i1);
    for (int ib = i0 + warp * 32; ib < i1; ib += 32 * kBinWarps) {
        const BinBatch bb = load_batch(b, ib, i1, lane);
        const float inv_w = 1.0f / (float)bb.w;
        if (bb.cnt && !bb.big) {
            unsigned long long m = bb.mask;
            while (m) {
                const int t = __ffsll((long long)m) - 1;
                m &= m - 1;
                atomicAdd(&cnt[rect_tile(t, inv_w, (int)bb.w, bb.mnx, bb.mny, b.gx)], 1u);
            }
        }
        unsigned bigs = __ballot_sync(0xffffffffu, bb.big != 0u);
        while (bigs) {
            const int l = __ffs(bigs) - 1;
            bigs &= bigs - 1;
            const int n = (int)__shfl_sync(0xffffffffu, bb.cnt, l), w = (int)__shfl_sync(0xffffffffu, bb.w, l);
            const int mnx = __shfl_sync(0xffffffffu, bb.mnx, l), mny = __shfl_sync(0xffffffffu, bb.mny, l);
            const float iw = 1.0f / (float)w;
            for (int t = lane; t < n; t += 32) atomicAdd(&cnt[rect_tile(t, iw, w, mnx, mny, b.gx)], 1u);
        }
    }
    __syncthreads();
    uint32_t* row = table + (size_t)c * b.num_tiles;
    for (int t = threadIdx.x; t < b.num_tiles; t += blockDim.x) row[t] = cnt[t];
}

// 3b: exclusive scan down every tile column of the [chunks][tiles] table; totals[t] = instances of tile t.  A CTA owns
// 32 adjacent tiles (one 128-byte line per table row) and splits the chunk axis over its 8 warps: per-warp partial sums,
// an 8-entry scan in shared memory, then the in-place exclusive scan of each warp's segment.
__global__ void __launch_bounds__(256) bin_colscan_kernel(int num_tiles, int chunks, uint32_t* __restrict__ table,
                                                         uint32_t* __restrict__ totals) {
    __shared__ uint32_t seg_sum[8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
    const int per = (chunks + 7) / 8, c0 = wid * per, c1 = min(chunks, c0 + per);
    const bool ok = t < num_tiles;
    uint32_t sum = 0;
    if (ok)
        for (int c = c0; c < c1; c++) sum += table[(size_t)c * num_tiles + t];
    seg_sum[wid][lane] = sum;
    __syncthreads();
    uint32_t run = 0;
    for (int w = 0; w < wid; w++) run += seg_sum[w][lane];
    if (ok) {
        int c = c0;
        for (; c + 4 <= c1; c += 4) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = table[(size_t)(c + u) * num_tiles + t];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                table[(size_t)(c + u) * num_tiles + t] = run;
                run += v[u];
            }
        }
        for (; c < c1; c++) {
            const uint32_t v = table[(size_t)c * num_tiles + t];
            table[(size_t)c * num_tiles + t] = run;
            run += v;
        }
        if (wid == 7) totals[t] = run;
    }
}

// 3b': exclusive scan over the tile totals -> base[t] and the tile ranges (DSR identifyTileRanges,
// rasterizer_impl.cu:116-138: empty tiles keep (0, 0)).  One CTA.
__global__ void __launch_bounds__(1024) bin_tilebase_kernel(int num_tiles, const uint32_t* __restrict__ totals,
                                                           uint32_t* __restrict__ base, uint2* __restrict__ ranges,
                                                           int64_t capacity, uint32_t* __restrict__ overflow) {
    __shared__ uint32_t warp_sums[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (num_tiles + 1023) / 1024;
    const int t0 = tid * per, t1 = min(num_tiles, t0 + per);
    uint32_t mine = 0;
    for (int t = t0; t < t1; t++) mine += totals[t];
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += v;
        }
        warp_sums[lane] = wi - ws;  // exclusive
        if (lane == 31 && (int64_t)wi > capacity) *overflow = 1u;  // more instances than the list buffer holds
    }
    __syncthreads();
    uint32_t run = warp_sums[wid] + incl - mine;
    for (int t = t0; t < t1; t++) {
        const uint32_t n = totals[t];
        base[t] = run;
        ranges[t] = n ? make_uint2(run, run + n) : make_uint2(0u, 0u);
        run += n;
    }
}

// 3b'': table[c][t] += base[t]: the table then holds the ABSOLUTE list position of the first instance of (run, tile), so
// that the scatter kernels need one global load per instance instead of two.
__global__ void __launch_bounds__(256) bin_addbase_kernel(int num_tiles, int chunks, const uint32_t* __restrict__ base,
                                                         uint32_t* __restrict__ table) {
    const size_t n = (size_t)chunks * num_tiles;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        table[i] += __ldg(base + (i % (size_t)num_tiles));
}

// 3c: every list entry straight to its final position, in depth order.  One CTA of kScatWarps warps per chunk.
// The chunk's instances are produced IN DEPTH ORDER into a ring in shared memory and consumed kQuad = 4 * kScatThreads
// at a time ("quad": warp w holds instances [128 w, 128 w + 128) of the quad, lane l the four instances 32 s + l), ranked
// by the whole CTA at once -- no warp ever waits for another warp's turn:
//   position = start of (chunk, tile)                      table[c][t] (absolute after bin_addbase)  (global, read-only)
//            + instances of the tile in earlier quads      cur[t]        u16 in shared memory
//            + ... in earlier warps of this quad           cnt[t]: one byte per warp in a 32-bit word, summed with dp4a
//            + ... earlier in this warp                    the warp's byte before its own sub-round added to it, plus the
//                                                          rank among the lanes with the same tile (one ballot per bit
//                                                          of the tile id)
// Two block barriers per quad (512 instances): byte updates -> [barrier] -> read word + cursor -> [barrier] -> group
// leaders clear their warp's byte (no other warp writes it), the first instance of each tile advances the cursor.  Four
// independent instances per thread also give the footprint arithmetic its instruction-level parallelism.
// Instances are materialised per "super-batch" of kScatThreads Gaussians (thread i owns Gaussian i of the depth order):
// each thread walks the set bits of its own footprint mask into the ring at its block-scan offset; footprints of more
// than 64 tiles (every tile of the rectangle) are expanded into the ring by the whole CTA.  Footprint data (cull rectangle
// + conic) is staged per Gaussian in one of 2 x kScatThreads slots (two super-batches alive) and fetched by slot when
// the 8 per-block bits of an entry are evaluated.  The next super-batch's records (and the footprint data they point at)
// are prefetched into registers.
// Measured alternatives, all within 5% of each other at ~0.4 ms for 10 M instances (profiles/r2_ncu_bin_scatter.txt):
// shared-memory atomics for the per-warp counts (2 cycles per lane on scattered addresses), MATCH.ANY (~350 cycles per
// warp on 32 mostly distinct values), one round per barrier pair, warps owning disjoint tile subsets with private queues
// (20 warps per SM but 50% more instructions).  The kernel is bound by dependent shared-memory / conversion latency at
// 12 warps per SM (the tile tables take 6 bytes of shared memory per tile and CTA).
constexpr int kScatWarps = 4;
constexpr int kScatThreads = 32 * kScatWarps;
constexpr int kQuad = 4 * kScatThreads;   // instances ranked between two pairs of barriers
constexpr int kRing = 2048;               // ring of staged instances (tile | slot << 16); power of two, >= 2 * kQuad
constexpr int kFpStride = 16;             // words per staged Gaussian footprint

size_t scatter_smem_bytes(int num_tiles) {
    return align_up((size_t)num_tiles * 4, 16) + align_up((size_t)num_tiles * 2, 16) + (size_t)kRing * 4 +
           (size_t)2 * kScatThreads * kFpStride * 4;
}

__global__ void __launch_bounds__(kScatThreads, 3)
bin_scatter_kernel(const BinArgs b, const uint32_t* __restrict__ table, const uint32_t* __restrict__ base, int64_t capacity,
                   uint32_t* __restrict__ point_list) {
    extern __shared__ __align__(16) unsigned char bin_smem[];
    uint32_t* cnt = reinterpret_cast<uint32_t*>(bin_smem);                                                // [num_tiles]
    uint16_t* cur = reinterpret_cast<uint16_t*>(bin_smem + align_up((size_t)b.num_tiles * 4, 16));         // [num_tiles]
    uint32_t* ring = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(cur) + align_up((size_t)b.num_tiles * 2, 16));
    float* fp = reinterpret_cast<float*>(ring + kRing);                                          // [2 * threads][kFpStride]
    __shared__ uint32_t s_cincl[kScatThreads];
    __shared__ uint32_t s_warp_tot[kScatWarps];
    __shared__ uint32_t s_bigmask[kScatWarps];
    __shared__ int s_end, s_bigw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x;
    const uint32_t* row = table + (size_t)c * b.num_tiles;
    for (int t = tid; t < b.num_tiles; t += kScatThreads) { cnt[t] = 0u; cur[t] = 0; }
    int i0, i1;
    chunk_range(b, c, lane, i0, i1);
    __syncthreads();
    const unsigned below = (1u << lane) - 1u;
    const uint32_t flower = (1u << (8u * (uint32_t)warp)) - 1u;
    const float inv_gx = 1.0f / (float)b.gx;
    const int tile_bits = 32 - __clz(max(1, b.num_tiles - 1));
    uint8_t* cnt8 = reinterpret_cast<uint8_t*>(cnt) + warp;  // this warp's byte of every tile's word
    uint32_t head = 0, count = 0;                              // ring state (block-uniform)

    auto entry_of = [&](uint32_t slot, uint32_t tile) -> uint32_t {
        const float* f = fp + slot * kFpStride;
        const float4 cr = *reinterpret_cast<const float4*>(f), q0 = *reinterpret_cast<const float4*>(f + 4),
                     q1 = *reinterpret_cast<const float4*>(f + 8);
        const int tile_y = __float2int_rd(((float)tile + 0.5f) * inv_gx), tile_x = (int)tile - tile_y * b.gx;
        return make_entry(__float_as_uint(f[13]), tile_x, tile_y, b.packed, cr, q0, q1, f[12]);
    };

    // Ranks and writes the first n (<= kQuad) instances of the ring.  Called by every thread.
    auto process_quad = [&](uint32_t n) {
        bool has[4], lead[4];
        uint32_t tile[4], entry[4], rkw[4], start[4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const uint32_t k = (uint32_t)(warp * 128 + s * 32 + lane);
            has[s] = k < n;
            const uint32_t e = has[s] ? ring[(head + k) & (kRing - 1)] : 0u;
            tile[s] = e & 0xffffu;
            entry[s] = has[s] ? entry_of(e >> 16, tile[s]) : 0u;
            start[s] = has[s] ? __ldg(row + tile[s]) : 0u;  // absolute position of the first instance of (run, tile)
        }
        // lanes of the same sub-round with the same tile
        unsigned peers[4];
#pragma unroll
        for (int s = 0; s < 4; s++) peers[s] = __ballot_sync(0xffffffffu, has[s]);
        for (int bit = 0; bit < tile_bits; bit++) {
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const unsigned set = __ballot_sync(0xffffffffu, (tile[s] >> bit) & 1u);
                peers[s] &= ((tile[s] >> bit) & 1u) ? set : ~set;
            }
        }
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const uint32_t rk = (uint32_t)__popc(peers[s] & below);
            lead[s] = has[s] && rk == 0u;
            uint32_t old = 0u;
            if (lead[s]) {  // the warp's byte counts its instances of the tile in this quad so far
                old = cnt8[tile[s] * 4u];
                cnt8[tile[s] * 4u] = (uint8_t)(old + (uint32_t)__popc(peers[s]));
            }
            old = __shfl_sync(0xffffffffu, old, (__ffs(peers[s]) - 1) & 31);
            rkw[s] = old + rk;
            __syncwarp();
        }
        __syncthreads();
        uint32_t v[4], c16[4], before[4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            v[s] = has[s] ? cnt[tile[s]] : 0u;
            c16[s] = has[s] ? (uint32_t)cur[tile[s]] : 0u;
            before[s] = __dp4a(v[s] & flower, 0x01010101u, 0u);  // instances of the tile in earlier warps
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if (lead[s]) cnt8[tile[s] * 4u] = 0;  // only this warp ever writes this byte
            if (has[s]) {
                if ((before[s] | rkw[s]) == 0u) cur[tile[s]] = (uint16_t)(c16[s] + __dp4a(v[s], 0x01010101u, 0u));
                const uint32_t pos = start[s] + c16[s] + before[s] + rkw[s];
                if ((int64_t)pos < capacity) point_list[pos] = entry[s];
            }
        }
        head = (head + n) & (kRing - 1);
        count -= n;
    };

    struct Pre {
        uint4 rec;
        unsigned long long mask;
        float4 cr, q0, q1;
        float r2;
        bool valid;
    };
    // Prefetches are unconditional loads from clamped addresses: a select on the loaded value (what a guarded load
    // compiles to) would wait for the data right here instead of one / two super-batches later.  Records past the end
    // of the chunk are ignored through `valid`; footprints of Gaussians without instances are never read.
    auto load_rec = [&](int ib, Pre& p) {
        const int i = min(ib + tid, b.P - 1);
        p.valid = ib + tid < i1;
        p.rec = __ldg(b.rec + i);
        p.mask = __ldg(b.mask + i);
    };
    auto load_fp = [&](Pre& p) {
        const uint32_t g = p.valid ? p.rec.x : 0u;
        p.cr = __ldg(b.cull4 + g);
        const float4* q = b.cullq + (size_t)g * 3;
        p.q0 = __ldg(q); p.q1 = __ldg(q + 1); p.r2 = __ldg(q + 2).x;
    };

    const int n_sb = (i1 - i0 + kScatThreads - 1) / kScatThreads;
    Pre me, nxt;
    load_rec(i0, me);
    load_fp(me);
    load_rec(i0 + kScatThreads, nxt);
    for (int sb = 0; sb < n_sb; sb++) {
        const uint32_t cnt_g = me.valid ? (me.rec.y & 0x7fffffffu) : 0u;
        const bool big = cnt_g > 64u;
        const uint32_t small_cnt = big ? 0u : cnt_g;
        const int mnx = (int)(me.rec.z & 0xffffu), mny = (int)(me.rec.z >> 16), w = (int)me.rec.w;
        const unsigned long long my_mask = me.mask;
        const uint32_t slot = (uint32_t)((sb & 1) * kScatThreads + tid);
        const uint32_t old_left = count;  // instances of the previous super-batch still in the ring (< kQuad)
        {   // staged footprint record of this thread's Gaussian (read by slot in entry_of)
            float* f = fp + slot * kFpStride;
            *reinterpret_cast<float4*>(f) = me.cr;
            *reinterpret_cast<float4*>(f + 4) = me.q0;
            *reinterpret_cast<float4*>(f + 8) = me.q1;
            *reinterpret_cast<float4*>(f + 12) = make_float4(me.r2, __uint_as_float(me.rec.x), __uint_as_float(cnt_g),
                                                               __uint_as_float(me.rec.z));
        }
        // block-wide inclusive scan of the small-footprint counts
        uint32_t c_incl = small_cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, c_incl, d);
            if (lane >= d) c_incl += u;
        }
        const unsigned bigs_w = __ballot_sync(0xffffffffu, big);
        if (lane == 31) s_warp_tot[warp] = c_incl;
        if (lane == 0) s_bigmask[warp] = bigs_w;
        __syncthreads();
        bool any_big = false;
#pragma unroll
        for (int u = 0; u < kScatWarps; u++) {
            if (u < warp) c_incl += s_warp_tot[u];
            any_big |= s_bigmask[u] != 0u;
        }
        s_cincl[tid] = c_incl;
        __syncthreads();
        // Prefetch for the next two super-batches, issued HERE: the 16-byte gathers occupy the memory pipe for hundreds
        // of cycles, and the shuffles of the scan above queued behind them when they were issued first (13% of all
        // stall samples).
        me = nxt;
        load_fp(me);
        load_rec(i0 + (sb + 2) * kScatThreads, nxt);
        const uint32_t c_excl = c_incl - small_cnt;
        const uint32_t sb_total = s_cincl[kScatThreads - 1];
        uint32_t consumed = 0;  // instances ranked during this super-batch

        int cur_g = 0;
        while (cur_g < kScatThreads) {  // block-uniform
            const uint32_t seg_base = cur_g ? s_cincl[cur_g - 1] : 0u;
            const uint32_t room = (uint32_t)kRing - count;
            int end = kScatThreads;
            if (any_big || sb_total - seg_base > room) {
                // first Gaussian >= cur_g that is big or would overflow the ring
                if (tid == 0) s_end = kScatThreads;
                __syncthreads();
                if (tid >= cur_g && (big || c_incl - seg_base > room)) atomicMin(&s_end, tid);
                __syncthreads();
                end = s_end;
            }
            const uint32_t n = end > cur_g ? s_cincl[end - 1] - seg_base : 0u;
            if (n) {
                if (tid >= cur_g && tid < end && small_cnt) {
                    // tile t of the rectangle (row-major, t < 64, width <= 64): t / w == (t * (65536 / w + 1)) >> 16
                    const uint32_t rcpw = 65536u / (uint32_t)w + 1u;
                    const uint32_t tile0 = (uint32_t)(mny * b.gx + mnx), dgx = (uint32_t)(b.gx - w);
                    uint32_t o = head + count + (c_excl - seg_base);
                    unsigned long long m = my_mask;
                    while (m) {
                        const uint32_t t = (uint32_t)__ffsll((long long)m) - 1u;
                        m &= m - 1;
                        const uint32_t ty = (t * rcpw) >> 16;   // tile = (mny + ty) * gx + mnx + (t - ty * w)
                        ring[o++ & (kRing - 1)] = (tile0 + t + ty * dgx) | (slot << 16);
                    }
                }
                count += n;
            }
            bool walked_big = false;
            if (end < kScatThreads && ((s_bigmask[end >> 5] >> (end & 31)) & 1u)) {
                // a footprint of more than 64 tiles: every tile of its rectangle, row-major, expanded by the whole CTA
                const uint32_t bslot = (uint32_t)((sb & 1) * kScatThreads + end);
                const float* f = fp + bslot * kFpStride;
                const int nb = (int)__float_as_uint(f[14]);
                const uint32_t org = __float_as_uint(f[15]);
                const int bx = (int)(org & 0xffffu), by = (int)(org >> 16);
                walked_big = true;
                int t0 = 0;
                if (tid == end) s_bigw = w;  // rectangle width: published by the owner thread
                __syncthreads();
                const int rw = s_bigw;
                const float iw = 1.0f / (float)rw;
                while (t0 < nb) {
                    const int m = min(nb - t0, (int)((uint32_t)kRing - count));
                    for (int t = tid; t < m; t += kScatThreads)
                        ring[(head + count + (uint32_t)t) & (kRing - 1)] = rect_tile(t0 + t, iw, rw, bx, by, b.gx) | (bslot << 16);
                    count += (uint32_t)m;
                    t0 += m;
                    __syncthreads();
                    while (count >= (uint32_t)kQuad) { process_quad(kQuad); consumed += kQuad; }
                    __syncthreads();  // ring reads of the quads above vs the writes of the next slice
                }
                cur_g = end + 1;
            } else {
                cur_g = end;
            }
            if (!walked_big) {
                __syncthreads();
                while (count >= (uint32_t)kQuad) { process_quad(kQuad); consumed += kQuad; }
            }
        }
        // The slots of this parity are rewritten two super-batches from now: nothing older than this super-batch may stay.
        if (consumed < old_left) {
            __syncthreads();
            process_quad(count);
        }
        __syncthreads();  // fp / ring / s_cincl are rewritten by the next super-batch
    }
    if (count) process_quad(count);
}

// ---------------------------------------------------------------------------------------------------------
// 3c', experimental variant (ISR_BIN_UNORDERED=1): UNORDERED scatter + sort of the (run, tile) segments.
// ---------------------------------------------------------------------------------------------------------
// Within a run the final position of an instance is  start of (run, tile) + its slot among the run's instances of that
// tile.  The ordered kernel above computes the slot as an in-order rank (ballots, per-warp count bytes, two barriers per
// 512 instances, 12 warps per SM).  Here the slot is simply the old value of a shared-memory atomic on the tile's 16-bit
// cursor -- no ranking, no count table (2 bytes of shared memory per tile: 20 warps per SM) -- so the instances of one
// (run, tile) segment land in arbitrary order, while the segments themselves are where they belong (the table of step
// 3b).  bin_fixup_kernel then sorts every segment of two or more entries by depth rank (rank[] = inverse of the depth
// order, written by gather_tiles): one thread per (run, tile), typical length 2-4 (registers, sorting network); segments
// longer than 32 entries go to bin_fixup_long_kernel (one CTA each, rank by counting through two scratch arrays).  The
// result is the same list as the ordered kernel's, entry for entry.
// (Flagging only the segments that can actually be out of order -- two instances of one tile between two block barriers
//  -- was tried first: most segments were flagged at cfg3, and the flagging (two dependent global atomics per conflict)
//  cost 25% of the kernel: profiles/r2_ncu_bin_scatter.txt.)
constexpr int kUThreads = 128;
constexpr int kStageCapU = 2048;          // staged instances (tile | owner thread << 16)

size_t scatter_u_smem_bytes(int num_tiles) {
    return align_up((size_t)((num_tiles + 1) / 2) * 4, 16) + (size_t)kStageCapU * 4 + (size_t)kUThreads * kFpStride * 4;
}

__global__ void __launch_bounds__(kUThreads, 5)
bin_scatter_unordered_kernel(const BinArgs b, const uint32_t* __restrict__ table, const uint32_t* __restrict__ base,
                             int64_t capacity, uint32_t* __restrict__ point_list) {
    extern __shared__ __align__(16) unsigned char bin_smem[];
    uint32_t* cur = reinterpret_cast<uint32_t*>(bin_smem);  // two 16-bit cursors per word: tile t -> word t/2, half t%2
    uint32_t* stage = reinterpret_cast<uint32_t*>(bin_smem + align_up((size_t)((b.num_tiles + 1) / 2) * 4, 16));
    float* fp = reinterpret_cast<float*>(stage + kStageCapU);  // [threads][kFpStride]
    __shared__ uint32_t s_cincl[kUThreads];
    __shared__ uint32_t s_warp_tot[kUThreads / 32];
    __shared__ uint32_t s_bigmask[kUThreads / 32];
    __shared__ int s_end, s_bigw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x;
    const uint32_t* row = table + (size_t)c * b.num_tiles;
    for (int t = tid; t < (b.num_tiles + 1) / 2; t += kUThreads) cur[t] = 0u;
    int i0, i1;
    chunk_range(b, c, lane, i0, i1);
    __syncthreads();
    const float inv_gx = 1.0f / (float)b.gx;

    auto entry_of = [&](uint32_t slot, uint32_t tile) -> uint32_t {
        const float* f = fp + slot * kFpStride;
        const float4 cr = *reinterpret_cast<const float4*>(f), q0 = *reinterpret_cast<const float4*>(f + 4),
                     q1 = *reinterpret_cast<const float4*>(f + 8);
        const int tile_y = __float2int_rd(((float)tile + 0.5f) * inv_gx), tile_x = (int)tile - tile_y * b.gx;
        return make_entry(__float_as_uint(f[13]), tile_x, tile_y, b.packed, cr, q0, q1, f[12]);
    };
    // The n staged instances take their slots (any order) and are written.  Called by every thread; begins with a
    // barrier (staging list complete); the caller puts one before the list is rewritten.
    auto process_batch = [&](uint32_t n) {
        __syncthreads();
        for (uint32_t k0 = 0; k0 < n; k0 += 4 * kUThreads) {
            uint32_t e[4], entry[4], st_base[4], st_row[4];
            bool has[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t k = k0 + (uint32_t)(u * kUThreads + tid);
                has[u] = k < n;
                e[u] = has[u] ? stage[k] : 0u;
                const uint32_t tile = e[u] & 0xffffu;
                // start of (run, tile): two L2 hits, consumed only at the store below (after footprint bits and atomic)
                st_base[u] = 0u;
                st_row[u] = has[u] ? __ldg(row + tile) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) entry[u] = has[u] ? entry_of(e[u] >> 16, e[u] & 0xffffu) : 0u;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (has[u]) {
                    const uint32_t tile = e[u] & 0xffffu, sh = (tile & 1u) * 16u;
                    const uint32_t slot = (atomicAdd(&cur[tile >> 1], 1u << sh) >> sh) & 0xffffu;
                    const uint32_t pos = st_base[u] + st_row[u] + slot;
                    if ((int64_t)pos < capacity) point_list[pos] = entry[u];
                }
            }
        }
    };

    struct Pre {
        uint4 rec;
        unsigned long long mask;
        float4 cr, q0, q1;
        float r2;
        bool valid;
    };
    auto load_rec = [&](int ib, Pre& p) {  // unconditional loads from clamped addresses (see the ordered kernel)
        const int i = min(ib + tid, b.P - 1);
        p.valid = ib + tid < i1;
        p.rec = __ldg(b.rec + i);
        p.mask = __ldg(b.mask + i);
    };
    auto load_fp = [&](Pre& p) {
        const uint32_t g = p.valid ? p.rec.x : 0u;
        p.cr = __ldg(b.cull4 + g);
        const float4* q = b.cullq + (size_t)g * 3;
        p.q0 = __ldg(q); p.q1 = __ldg(q + 1); p.r2 = __ldg(q + 2).x;
    };

    const int n_sb = (i1 - i0 + kUThreads - 1) / kUThreads;
    Pre me, nxt;
    load_rec(i0, me);
    load_fp(me);
    load_rec(i0 + kUThreads, nxt);
    for (int sb = 0; sb < n_sb; sb++) {
        const uint32_t cnt_g = me.valid ? (me.rec.y & 0x7fffffffu) : 0u;
        const bool big = cnt_g > 64u;
        const uint32_t small_cnt = big ? 0u : cnt_g;
        const int mnx = (int)(me.rec.z & 0xffffu), mny = (int)(me.rec.z >> 16), w = (int)me.rec.w;
        const unsigned long long my_mask = me.mask;
        {   // staged footprint record of this thread's Gaussian (read by owner index in entry_of)
            float* f = fp + tid * kFpStride;
            *reinterpret_cast<float4*>(f) = me.cr;
            *reinterpret_cast<float4*>(f + 4) = me.q0;
            *reinterpret_cast<float4*>(f + 8) = me.q1;
            *reinterpret_cast<float4*>(f + 12) = make_float4(me.r2, __uint_as_float(me.rec.x), __uint_as_float(cnt_g),
                                                               __uint_as_float(me.rec.z));
        }
        uint32_t c_incl = small_cnt;  // block-wide inclusive scan of the small-footprint counts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, c_incl, d);
            if (lane >= d) c_incl += u;
        }
        const unsigned bigs_w = __ballot_sync(0xffffffffu, big);
        if (lane == 31) s_warp_tot[warp] = c_incl;
        if (lane == 0) s_bigmask[warp] = bigs_w;
        __syncthreads();
        bool any_big = false;
#pragma unroll
        for (int u = 0; u < kUThreads / 32; u++) {
            if (u < warp) c_incl += s_warp_tot[u];
            any_big |= s_bigmask[u] != 0u;
        }
        s_cincl[tid] = c_incl;
        __syncthreads();
        me = nxt;  // prefetch for the next two super-batches (issued after the scan, see the ordered kernel)
        load_fp(me);
        load_rec(i0 + (sb + 2) * kUThreads, nxt);
        const uint32_t c_excl = c_incl - small_cnt;
        const uint32_t sb_total = s_cincl[kUThreads - 1];
        const bool simple = !any_big && sb_total <= (uint32_t)kStageCapU;

        int cur_g = 0;
        bool first = true;
        while (cur_g < kUThreads) {  // block-uniform
            const uint32_t seg_base = cur_g ? s_cincl[cur_g - 1] : 0u;
            int end = kUThreads;
            if (!simple) {  // first Gaussian >= cur_g that is big or would overflow the staging list
                if (tid == 0) s_end = kUThreads;
                __syncthreads();  // (also: the previous segment's staging list has been consumed)
                if (tid >= cur_g && (big || c_incl - seg_base > (uint32_t)kStageCapU)) atomicMin(&s_end, tid);
                __syncthreads();
                end = s_end;
            } else if (!first) {
                break;
            }
            first = false;
            const uint32_t n = end > cur_g ? s_cincl[end - 1] - seg_base : 0u;
            if (n) {
                if (tid >= cur_g && tid < end && small_cnt) {
                    // tile t of the rectangle (row-major, t < 64, width <= 64): t / w == (t * (65536 / w + 1)) >> 16
                    const uint32_t rcpw = 65536u / (uint32_t)w + 1u;
                    const uint32_t tile0 = (uint32_t)(mny * b.gx + mnx), dgx = (uint32_t)(b.gx - w);
                    uint32_t o = c_excl - seg_base;
                    unsigned long long m = my_mask;
                    while (m) {
                        const uint32_t t = (uint32_t)__ffsll((long long)m) - 1u;
                        m &= m - 1;
                        const uint32_t ty = (t * rcpw) >> 16;
                        stage[o++] = (tile0 + t + ty * dgx) | ((uint32_t)tid << 16);
                    }
                }
                process_batch(n);
            }
            if (end < kUThreads && ((s_bigmask[end >> 5] >> (end & 31)) & 1u)) {
                // a footprint of more than 64 tiles: every tile of its rectangle, row-major, expanded by the whole CTA
                const float* f = fp + end * kFpStride;
                const int nb = (int)__float_as_uint(f[14]);
                const uint32_t org = __float_as_uint(f[15]);
                const int bx = (int)(org & 0xffffu), by = (int)(org >> 16);
                if (tid == end) s_bigw = w;  // rectangle width: published by the owner thread
                for (int t0 = 0; t0 < nb; t0 += kStageCapU) {
                    __syncthreads();  // s_bigw visible; the previous staging list has been consumed
                    const int rw = s_bigw;
                    const float iw = 1.0f / (float)rw;
                    const int m = min(nb - t0, kStageCapU);
                    for (int t = tid; t < m; t += kUThreads) stage[t] = rect_tile(t0 + t, iw, rw, bx, by, b.gx) | ((uint32_t)end << 16);
                    process_batch((uint32_t)m);
                }
                cur_g = end + 1;
            } else {
                cur_g = end;
            }
        }
        __syncthreads();  // fp / staging list / s_cincl are rewritten by the next super-batch
    }
}

// Sorts every (run, tile) segment of two or more entries by depth rank.  One thread per segment (t fastest: the table
// rows are read coalesced); up to 4 entries in registers (all loads in flight together, 5-comparator network), up to 32
// by insertion in local memory, longer ones are queued for bin_fixup_long_kernel.  (A warp-cooperative version -- 32
// segments gathered into shared memory, rank by counting with every lane busy -- was measured slower: 150 vs 111 us.)
__global__ void __launch_bounds__(256)
bin_fixup_kernel(int num_tiles, int chunks, const uint32_t* __restrict__ table, const uint32_t* __restrict__ totals,
                 const uint32_t* __restrict__ base, const uint32_t* __restrict__ rank, int packed, int64_t capacity,
                 uint32_t* __restrict__ fix_counters, uint32_t* __restrict__ long_list, uint32_t long_cap,
                 uint32_t* __restrict__ point_list) {
    const uint32_t idm = packed ? kIdMask : 0xffffffffu;
    const size_t n_seg = (size_t)chunks * num_tiles;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_seg; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i / (size_t)num_tiles), t = (uint32_t)(i - (size_t)c * num_tiles);
        const uint32_t s0 = table[i];
        const uint32_t s1 = (int)c + 1 < chunks ? table[i + num_tiles] : base[t] + totals[t];  // (the table is absolute)
        const uint32_t len = s1 - s0;
        if (len < 2u) continue;
        const uint32_t off = s0;
        if ((int64_t)off + len > capacity) continue;  // (list buffer too small: the host repeats the binning)
        uint32_t* p = point_list + off;
        if (len <= 4u) {
            uint32_t e[4], k[4];
#pragma unroll
            for (int j = 0; j < 4; j++) e[j] = (uint32_t)j < len ? p[j] : 0u;
#pragma unroll
            for (int j = 0; j < 4; j++) k[j] = (uint32_t)j < len ? __ldg(rank + (e[j] & idm)) : 0xffffffffu;
            bool moved = false;
            auto cx = [&](int a, int bb) {
                if (k[a] > k[bb]) { const uint32_t tk = k[a], te = e[a]; k[a] = k[bb]; e[a] = e[bb]; k[bb] = tk; e[bb] = te; moved = true; }
            };
            cx(0, 1); cx(2, 3); cx(0, 2); cx(1, 3); cx(1, 2);
            if (moved) {
#pragma unroll
                for (int j = 0; j < 4; j++) if ((uint32_t)j < len) p[j] = e[j];
            }
            continue;
        }
        if (len > 32u) {
            const uint32_t at = atomicAdd(fix_counters + 1, 1u);
            if (at < long_cap) long_list[at] = (uint32_t)i;
            continue;
        }
        uint32_t e[32], k[32];
        for (uint32_t j = 0; j < len; j++) e[j] = p[j];
        for (uint32_t j = 0; j < len; j++) k[j] = __ldg(rank + (e[j] & idm));
        for (uint32_t j = 1; j < len; j++) {
            const uint32_t ej = e[j], kj = k[j];
            uint32_t q = j;
            while (q > 0 && k[q - 1] > kj) { e[q] = e[q - 1]; k[q] = k[q - 1]; q--; }
            e[q] = ej; k[q] = kj;
        }
        for (uint32_t j = 0; j < len; j++) p[j] = e[j];
    }
}

// One CTA per long segment: position of an entry = number of entries of the segment with a smaller depth rank.
__global__ void __launch_bounds__(256)
bin_fixup_long_kernel(int num_tiles, int chunks, const uint32_t* __restrict__ table, const uint32_t* __restrict__ totals,
                      const uint32_t* __restrict__ base, const uint32_t* __restrict__ rank, int packed,
                      const uint32_t* __restrict__ long_list, uint32_t long_cap, const uint32_t* __restrict__ fix_counters,
                      uint32_t* __restrict__ scratch_k, uint32_t* __restrict__ scratch_e, uint32_t* __restrict__ point_list) {
    const uint32_t n_long = min(fix_counters[1], long_cap);
    const uint32_t idm = packed ? kIdMask : 0xffffffffu;
    for (uint32_t i = blockIdx.x; i < n_long; i += gridDim.x) {
        const uint32_t bit = long_list[i];
        const uint32_t c = bit / (uint32_t)num_tiles, t = bit - c * (uint32_t)num_tiles;
        const uint32_t s0 = table[(size_t)c * num_tiles + t];
        const uint32_t s1 = (int)c + 1 < chunks ? table[(size_t)(c + 1) * num_tiles + t] : base[t] + totals[t];
        const uint32_t len = s1 - s0, off = s0;
        uint32_t* p = point_list + off;
        uint32_t* sk = scratch_k + off;
        uint32_t* se = scratch_e + off;
        for (uint32_t j = threadIdx.x; j < len; j += blockDim.x) {
            const uint32_t ej = p[j];
            se[j] = ej;
            sk[j] = __ldg(rank + (ej & idm));
        }
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < len; j += blockDim.x) {
            const uint32_t kj = sk[j];
            uint32_t pos = 0;
            for (uint32_t q = 0; q < len; q++) pos += sk[q] < kj ? 1u : 0u;  // ranks are distinct
            p[pos] = se[j];
        }
        __syncthreads();
    }
}

// Phase B head: stable partition of the instances by tile (fused with emission) + tile ranges.  `R` is the CAPACITY of
// the list buffer (>= the emitted instance count, which the kernels read from device memory).
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
    if (R <= 0 || P <= 0) {
        ISR_CUDA_TRY(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)num_tiles, stream));
        return ISR_OK;
    }
    uint32_t* point_list = reinterpret_cast<uint32_t*>(b + bl.point_list);
    const int packed = entries_packed(P) ? 1 : 0;
    const BinChunks bc = bin_chunks(num_tiles, R);
    if (bc.chunks > 0) {
        BinArgs ba;
        ba.P = P; ba.gx = gx; ba.gy = gy; ba.W = a.W; ba.H = a.H; ba.num_tiles = num_tiles; ba.chunks = bc.chunks; ba.packed = packed;
        ba.offsets = reinterpret_cast<const uint32_t*>(g + gl.offsets);
        ba.rec = reinterpret_cast<const uint4*>(g + gl.bin_rec);
        ba.mask = reinterpret_cast<const unsigned long long*>(g + gl.bin_mask);
        ba.cull4 = reinterpret_cast<const float4*>(g + gl.cull);
        ba.cullq = reinterpret_cast<const float4*>(g + gl.cullq);
        uint32_t* table = reinterpret_cast<uint32_t*>(b + bl.table);
        uint32_t* totals = reinterpret_cast<uint32_t*>(b + bl.totals);
        uint32_t* base = reinterpret_cast<uint32_t*>(b + bl.base);
        uint32_t* overflow = reinterpret_cast<uint32_t*>(g + gl.counters) + 5;
        ISR_CUDA_TRY(cudaFuncSetAttribute(bin_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bc.smem_count));
        ISR_CUDA_TRY(cudaFuncSetAttribute(bin_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bc.smem_scatter));
        bin_count_kernel<<<bc.chunks, 32 * kBinWarps, bc.smem_count, stream>>>(ba, table); note_launch();
        ISR_CUDA_TRY(cudaGetLastError());
        bin_colscan_kernel<<<(num_tiles + 31) / 32, 256, 0, stream>>>(num_tiles, bc.chunks, table, totals); note_launch();
        ISR_CUDA_TRY(cudaGetLastError());
        bin_tilebase_kernel<<<1, 1024, 0, stream>>>(num_tiles, totals, base, ranges, R, overflow); note_launch();
        ISR_CUDA_TRY(cudaGetLastError());
        bin_addbase_kernel<<<bin_sm_count() * 8, 256, 0, stream>>>(num_tiles, bc.chunks, base, table); note_launch();
        ISR_CUDA_TRY(cudaGetLastError());
        if (!bc.unordered) {
            bin_scatter_kernel<<<bc.chunks, kScatThreads, bc.smem_scatter, stream>>>(ba, table, base, R, point_list); note_launch();
            ISR_CUDA_TRY(cudaGetLastError());
            return ISR_OK;
        }
        uint32_t* fix_counters = reinterpret_cast<uint32_t*>(b + bl.fix_counters);
        uint32_t* long_list = reinterpret_cast<uint32_t*>(b + bl.long_list);
        ISR_CUDA_TRY(cudaMemsetAsync(fix_counters, 0, 256, stream));
        ISR_CUDA_TRY(cudaFuncSetAttribute(bin_scatter_unordered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bc.smem_scatter));
        bin_scatter_unordered_kernel<<<bc.chunks, kUThreads, bc.smem_scatter, stream>>>(ba, table, base, R, point_list); note_launch();
        ISR_CUDA_TRY(cudaGetLastError());
        const uint32_t* rank = reinterpret_cast<const uint32_t*>(g + gl.rank);
        bin_fixup_kernel<<<bin_sm_count() * 16, 256, 0, stream>>>(num_tiles, bc.chunks, table, totals, base, rank, packed, R,
                                                                 fix_counters, long_list, (uint32_t)bl.long_cap, point_list); note_launch();
        ISR_CUDA_TRY(cudaGetLastError());
        bin_fixup_long_kernel<<<bin_sm_count(), 256, 0, stream>>>(num_tiles, bc.chunks, table, totals, base, rank, packed, long_list,
                                                                 (uint32_t)bl.long_cap, fix_counters,
                                                                 reinterpret_cast<uint32_t*>(b + bl.scratch_k),
                                                                 reinterpret_cast<uint32_t*>(b + bl.scratch_e), point_list); note_launch();
        ISR_CUDA_TRY(cudaGetLastError());
        return ISR_OK;
    }
    // ---- fallback: (tile id, entry) pairs + radix sort by tile id (CUB) + tile ranges --------------------------------
    ISR_CUDA_TRY(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)num_tiles, stream));
    uint32_t* point_list_alt = reinterpret_cast<uint32_t*>(b + bl.point_list_alt);
    uint32_t* tile_keys = reinterpret_cast<uint32_t*>(b + bl.tile_keys);
    uint32_t* tile_keys_alt = reinterpret_cast<uint32_t*>(b + bl.tile_keys_alt);
    uint32_t* big_count = reinterpret_cast<uint32_t*>(g + gl.counters) + 4;
    uint2* big_list = reinterpret_cast<uint2*>(g + gl.big_list);
    ISR_CUDA_TRY(cudaMemsetAsync(big_count, 0, sizeof(uint32_t), stream));
    emit_instances_kernel<<<(P + 255) / 256, 256, 0, stream>>>(
        P, reinterpret_cast<const uint32_t*>(g + gl.order), reinterpret_cast<const uint32_t*>(g + gl.offsets), a.radii,
        reinterpret_cast<const Splat*>(g + gl.splat), reinterpret_cast<const uint4*>(g + gl.tfoot),
        reinterpret_cast<const uint32_t*>(g + gl.tcount), reinterpret_cast<const float4*>(g + gl.cull),
        reinterpret_cast<const float4*>(g + gl.cullq), gx, gy, a.W, a.H, packed, tile_keys_alt, point_list_alt, big_count,
        big_list); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    emit_big_kernel<<<592, 256, 0, stream>>>(big_count, big_list, a.radii, reinterpret_cast<const Splat*>(g + gl.splat),
                                             reinterpret_cast<const float4*>(g + gl.cull),
                                             reinterpret_cast<const float4*>(g + gl.cullq), gx, gy, a.W, a.H, packed,
                                             tile_keys_alt, point_list_alt); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    int bits = 1;
    while ((1 << bits) < num_tiles) bits++;
    size_t temp_bytes = bl.temp_bytes;
    ISR_CUDA_TRY(cub::DeviceRadixSort::SortPairs(b + bl.temp, temp_bytes, tile_keys_alt, tile_keys, point_list_alt,
                                                 point_list, (int)R, 0, bits, stream));
    tile_ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(R, tile_keys, ranges); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

}  // namespace isr
