// isr_knn.cu -- distCUDA2 replacement: mean squared distance to the 3 nearest neighbours of every point.
// Reference: submodules/simple-knn/simple_knn.cu:186-222 (Morton sort + 1024-point boxes + pruning; two
// blocking D2H copies and cudaMalloc/cudaFree per call).  Same exact result, different structure: points are
// bucketed into a uniform grid (~4 points per cell) with one radix sort, and every point grows a cube of
// cells ring by ring until its third-best distance is provably final.  No host round trip, no allocation.
#include <cfloat>
#include <cub/cub.cuh>

#include "isr_common.cuh"

namespace isr {

struct KnnWs {
    size_t bbox, keys, keys_alt, ids, ids_alt, cell_start, temp, temp_bytes, total;
    int G;
    explicit KnnWs(int P) {
        G = 1;
        while ((int64_t)G * G * G * 4 < (int64_t)P && G < 256) G++;
        const size_t p = (size_t)(P > 0 ? P : 1);
        const size_t cells = (size_t)G * G * G;
        size_t o = 0;
        bbox = o;       o = align_up(o + 8 * 4, 256);
        keys = o;       o = align_up(o + p * 4, 256);
        keys_alt = o;   o = align_up(o + p * 4, 256);
        ids = o;        o = align_up(o + p * 4, 256);
        ids_alt = o;    o = align_up(o + p * 4, 256);
        cell_start = o; o = align_up(o + (cells + 1) * 4, 256);
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                        (uint32_t*)nullptr, (int)p, 0, 32);
        temp_bytes = tb + 256;
        temp = o;       o = align_up(o + temp_bytes, 256);
        total = o;
    }
};

__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void knn_bbox_init(int* bbox) {
    if (threadIdx.x < 3) bbox[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) bbox[threadIdx.x] = (int)0x80000000;
}

__global__ void knn_bbox_kernel(int P, const float* __restrict__ pts, int* __restrict__ bbox) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
        for (int c = 0; c < 3; c++) {
            const float v = pts[3 * (size_t)i + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    for (int c = 0; c < 3; c++) {
        for (int off = 16; off > 0; off >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], off));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], off));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bbox + c, float_to_ordered(mn[c]));
            atomicMax(bbox + 3 + c, float_to_ordered(mx[c]));
        }
    }
}

__device__ __forceinline__ void knn_cell_of(const float* p, const int* bbox, int G, int& ix, int& iy, int& iz,
                                            float* origin, float* cell) {
    for (int c = 0; c < 3; c++) {
        const float lo = ordered_to_float(bbox[c]), hi = ordered_to_float(bbox[3 + c]);
        origin[c] = lo;
        cell[c] = fmaxf((hi - lo) / (float)G, 1e-30f);
    }
    ix = min(G - 1, max(0, (int)((p[0] - origin[0]) / cell[0])));
    iy = min(G - 1, max(0, (int)((p[1] - origin[1]) / cell[1])));
    iz = min(G - 1, max(0, (int)((p[2] - origin[2]) / cell[2])));
}

__global__ void knn_keys_kernel(int P, const float* __restrict__ pts, const int* __restrict__ bbox, int G,
                                uint32_t* __restrict__ keys, uint32_t* __restrict__ ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float p[3] = {pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]};
    int ix, iy, iz;
    float origin[3], cell[3];
    knn_cell_of(p, bbox, G, ix, iy, iz, origin, cell);
    keys[i] = (uint32_t)(ix + G * (iy + G * iz));
    ids[i] = (uint32_t)i;
}

// cell_start[c] = first sorted position whose key >= c  (cell_start[cells] = P)
__global__ void knn_cell_start_kernel(int P, int cells, const uint32_t* __restrict__ keys_sorted,
                                      uint32_t* __restrict__ cell_start) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > P) return;
    const int cur = (i < P) ? (int)keys_sorted[i] : cells;
    const int prev = (i == 0) ? -1 : (int)keys_sorted[i - 1];
    for (int c = prev + 1; c <= cur; c++) cell_start[c] = (uint32_t)i;
}

__device__ __forceinline__ void knn_update(float dist, float* best) {
#pragma unroll
    for (int j = 0; j < 3; j++)
        if (best[j] > dist) { const float t = best[j]; best[j] = dist; dist = t; }
}

__global__ void __launch_bounds__(128)
knn_search_kernel(int P, const float* __restrict__ pts, const int* __restrict__ bbox, int G,
                  const uint32_t* __restrict__ ids_sorted, const uint32_t* __restrict__ cell_start,
                  float* __restrict__ out) {
    const int sidx = blockIdx.x * blockDim.x + threadIdx.x;  // walk points in cell order for locality
    if (sidx >= P) return;
    const int i = (int)ids_sorted[sidx];
    const float p[3] = {pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]};
    int cx, cy, cz;
    float origin[3], cell[3];
    knn_cell_of(p, bbox, G, cx, cy, cz, origin, cell);
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    auto visit = [&](int x, int y, int z) {
        const int c = x + G * (y + G * z);
        const uint32_t b = cell_start[c], e = cell_start[c + 1];
        for (uint32_t k = b; k < e; k++) {
            const int j = (int)ids_sorted[k];
            if (j == i) continue;
            const float dx = pts[3 * (size_t)j] - p[0], dy = pts[3 * (size_t)j + 1] - p[1],
                        dz = pts[3 * (size_t)j + 2] - p[2];
            // (dx*dx + dy*dy) + dz*dz as nvcc contracts it in simple_knn.cu:123-124
            knn_update(fma_(dz, dz, fma_(dx, dx, mul(dy, dy))), best);
        }
    };
    for (int r = 0; r < G; r++) {
        // visit the shell of the cube of cells [c-r, c+r] (clipped to the grid)
        const int x0 = cx - r, x1 = cx + r, y0 = cy - r, y1 = cy + r, z0 = cz - r, z1 = cz + r;
        for (int z = max(z0, 0); z <= min(z1, G - 1); z++)
            for (int y = max(y0, 0); y <= min(y1, G - 1); y++) {
                const bool full_row = (z == z0 || z == z1 || y == y0 || y == y1);
                if (full_row) {
                    for (int x = max(x0, 0); x <= min(x1, G - 1); x++) visit(x, y, z);
                } else {
                    if (x0 >= 0) visit(x0, y, z);
                    if (x1 != x0 && x1 <= G - 1) visit(x1, y, z);
                }
            }
        // every unvisited point lies outside the cube of cells [c-r, c+r]: lower bound on its distance.  An axis along
        // which the cloud has no extent (planar / collinear input: every point sits in cell 0 of that axis) is covered
        // from the first ring on -- without this the bound on that axis stays ~0 and a planar cloud scans all G rings.
        float bound = FLT_MAX;
        bool covers_all = true;
        const int lo[3] = {x0, y0, z0}, hi[3] = {x1, y1, z1};
        for (int a = 0; a < 3; a++) {
            if (!(ordered_to_float(bbox[3 + a]) > origin[a])) continue;
            if (lo[a] > 0) { covers_all = false; bound = fminf(bound, p[a] - (origin[a] + cell[a] * (float)lo[a])); }
            if (hi[a] < G - 1) { covers_all = false; bound = fminf(bound, (origin[a] + cell[a] * (float)(hi[a] + 1)) - p[a]); }
        }
        if (covers_all) break;
        bound = fmaxf(bound, 0.0f) * 0.999f;  // slack for the fp32 rounding of the cell faces
        if (best[2] <= bound * bound) break;
    }
    out[i] = __fdiv_rn(add(add(best[0], best[1]), best[2]), 3.0f);
}

size_t knn_ws_bytes(int P) { return KnnWs(P).total; }

int launch_knn(int P, const float* points, float* out, void* ws, size_t ws_bytes, cudaStream_t stream) {
    (void)ws_bytes;
    KnnWs L(P);
    char* w = static_cast<char*>(ws);
    int* bbox = reinterpret_cast<int*>(w + L.bbox);
    uint32_t* keys = reinterpret_cast<uint32_t*>(w + L.keys);
    uint32_t* keys_alt = reinterpret_cast<uint32_t*>(w + L.keys_alt);
    uint32_t* ids = reinterpret_cast<uint32_t*>(w + L.ids);
    uint32_t* ids_alt = reinterpret_cast<uint32_t*>(w + L.ids_alt);
    uint32_t* cell_start = reinterpret_cast<uint32_t*>(w + L.cell_start);
    const int G = L.G, cells = G * G * G;
    knn_bbox_init<<<1, 32, 0, stream>>>(bbox); note_launch();
    knn_bbox_kernel<<<148 * 4, 256, 0, stream>>>(P, points, bbox); note_launch();
    knn_keys_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, points, bbox, G, keys, ids); note_launch();
    int bits = 1;
    while ((1 << bits) < cells) bits++;
    size_t tb = L.temp_bytes;
    ISR_CUDA_TRY(cub::DeviceRadixSort::SortPairs(w + L.temp, tb, keys, keys_alt, ids, ids_alt, P, 0, bits, stream));
    knn_cell_start_kernel<<<(P + 1 + 255) / 256, 256, 0, stream>>>(P, cells, keys_alt, cell_start); note_launch();
    knn_search_kernel<<<(P + 127) / 128, 128, 0, stream>>>(P, points, bbox, G, ids_alt, cell_start, out); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

}  // namespace isr
