// isr_blend_fwd.cu -- K6: per-tile front-to-back alpha compositing of RGB + 7 auxiliary maps + F semantic
// feature channels + the (gaussian, pixel) pair list.   Reference: DSR/cuda_rasterizer/forward.cu:256-462.
//
// Mapping: one CTA (256 threads) per 16x16 tile, one thread per pixel, a warp covers an 8x4 pixel block.
// Per batch of 256 list entries every thread stages ONE instance into shared memory (splat record 64 B,
// cull rect 16 B, rgb 16 B, features 4F B -- all 16-byte vector loads), so the inner loop issues no global
// loads (the reference fetches rgb and features from global per contributing (pixel, Gaussian) pair).
// Each warp first tests 32 staged Gaussians at a time against its own 8x4 pixel block with the conservative
// cull rectangle computed in K1 (one lane per Gaussian, one ballot), and only walks the survivors in list
// order -- exact, because a culled Gaussian provably fails the alpha >= 1/255 test on every pixel of the block.
// Pair-list entries are staged per warp in shared memory and flushed with one global atomic per 32+ pairs
// (the reference does one global atomic per pair on a single counter).
#include "isr_common.cuh"

namespace isr {

constexpr int kBatch = 256;
constexpr int kPairStage = 96;  // per-warp staging slots (int2)

template <int FP>  // feature dim padded to a multiple of 4 (0, 4, 8, 16, 24, 32)
struct FwdSmem {
    static constexpr int kFeatVec = FP / 4;
    static constexpr size_t bytes = (size_t)kBatch * (64 + 16 + 16 + 4 * FP + 4) + 8 * kPairStage * 8;
};

template <int FP, bool kPairs>
__global__ void __launch_bounds__(256)
blend_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int F,
                 const float4* __restrict__ splats, const float4* __restrict__ cull4, const float4* __restrict__ rgb4,
                 const float* __restrict__ extras, const float* __restrict__ bg, float* __restrict__ final_T,
                 uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_others,
                 float* __restrict__ out_extra, int2* __restrict__ pairs, int64_t pair_cap, int* __restrict__ pair_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_splat = reinterpret_cast<float4*>(smem_raw);                 // [kBatch][4]
    float4* s_cull = s_splat + kBatch * 4;                                 // [kBatch]
    float4* s_rgb = s_cull + kBatch;                                       // [kBatch]
    float4* s_feat = s_rgb + kBatch;                                       // [kBatch][FP/4]
    int* s_id = reinterpret_cast<int*>(s_feat + kBatch * (FP / 4));        // [kBatch]
    int2* s_pairs = reinterpret_cast<int2*>(s_id + kBatch);                // [8][kPairStage]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (W + TILE - 1) / TILE;
    const int tile_id = blockIdx.y * tiles_x + blockIdx.x;
    // warp -> 8x4 block inside the tile
    const int wx0 = blockIdx.x * TILE + (warp & 1) * 8;
    const int wy0 = blockIdx.y * TILE + (warp >> 1) * 4;
    const int pxi = wx0 + (lane & 7), pyi = wy0 + (lane >> 3);
    const bool inside = pxi < W && pyi < H;
    const uint32_t pix_id = (uint32_t)W * (uint32_t)pyi + (uint32_t)pxi;
    const float pixx = (float)pxi, pixy = (float)pyi;
    // pixel block bounds of this warp (inclusive), clipped to the image
    const float bx0 = (float)wx0, by0 = (float)wy0;
    const float bx1 = (float)min(wx0 + 7, W - 1), by1 = (float)min(wy0 + 3, H - 1);

    const uint2 range = ranges[tile_id];
    const int n_total = (int)(range.y - range.x);
    const float c1 = __fdiv_rn(kFar, __fsub_rn(kFar, kNear));

    bool done = !inside;
    float T = 1.0f;
    float C0 = 0, C1 = 0, C2 = 0, N0 = 0, N1 = 0, N2 = 0, D = 0, M1 = 0, M2 = 0, dist = 0, median_depth = 0;
    uint32_t last_contributor = 0, median_contributor = 0;
    float E[FP > 0 ? FP : 1];
#pragma unroll
    for (int ch = 0; ch < FP; ch++) E[ch] = 0.0f;
    int wcount = 0;  // staged pairs of this warp (warp-uniform)
    int2* my_pairs = s_pairs + warp * kPairStage;

    for (int base = 0; base < n_total; base += kBatch) {
        // whole-block early exit (forward.cu:331-333)
        if (__syncthreads_count(done) == 256) break;
        const int n_batch = min(kBatch, n_total - base);
        if (tid < n_batch) {
            const int g = (int)point_list[range.x + base + tid];
            s_id[tid] = g;
            const float4* sp = splats + (size_t)g * 4;
            s_splat[tid * 4 + 0] = __ldg(sp + 0);
            s_splat[tid * 4 + 1] = __ldg(sp + 1);
            s_splat[tid * 4 + 2] = __ldg(sp + 2);
            s_splat[tid * 4 + 3] = __ldg(sp + 3);
            s_cull[tid] = __ldg(cull4 + g);
            s_rgb[tid] = __ldg(rgb4 + g);
            if (FP > 0) {
                if ((F & 3) == 0) {
                    const float4* fp = reinterpret_cast<const float4*>(extras + (size_t)g * F);
#pragma unroll
                    for (int v = 0; v < FP / 4; v++)
                        s_feat[tid * (FP / 4) + v] = (v * 4 < F) ? __ldg(fp + v) : make_float4(0, 0, 0, 0);
                } else {
                    float* dstf = reinterpret_cast<float*>(s_feat + tid * (FP / 4));
#pragma unroll
                    for (int ch = 0; ch < FP; ch++) dstf[ch] = (ch < F) ? __ldg(extras + (size_t)g * F + ch) : 0.0f;
                }
            }
        }
        __syncthreads();

        for (int j0 = 0; j0 < n_batch; j0 += 32) {
            if (__all_sync(0xffffffffu, done)) break;
            // one lane per staged Gaussian: does its cull rect overlap this warp's pixel block?
            bool ov = false;
            if (j0 + lane < n_batch) {
                const float4 cr = s_cull[j0 + lane];
                ov = !(cr.z < bx0 || cr.x > bx1 || cr.w < by0 || cr.y > by1);
            }
            unsigned todo = __ballot_sync(0xffffffffu, ov);
            while (todo) {
                const int j = j0 + __ffs(todo) - 1;
                todo &= todo - 1;
                bool hit = false;
                float w = 0.0f;
                if (!done) {
                    const float* s = reinterpret_cast<const float*>(s_splat + j * 4);
                    PairEval e;
                    if (eval_pair<false>(pixx, pixy, s, e)) {
                        const float test_T = mul(T, sub(1.0f, e.alpha));
                        if (test_T < kTMin) {
                            done = true;
                        } else {
                            hit = true;
                            const uint32_t contributor = (uint32_t)(base + j + 1);
                            w = mul(e.alpha, T);
                            const float A = sub(1.0f, T);
                            const float m = mul(c1, sub(1.0f, mul(kNear, rcp(e.depth))));
                            const float mm = mul(m, m);
                            const float dt = fma_(-add(m, m), M1, fma_(mm, A, M2));
                            dist = fma_(dt, w, dist);
                            D = fma_(e.depth, w, D);
                            M1 = fma_(m, w, M1);
                            M2 = fma_(mm, w, M2);
                            if (T > 0.5f) { median_depth = e.depth; median_contributor = contributor; }
                            N0 = fma_(s[11], w, N0); N1 = fma_(s[12], w, N1); N2 = fma_(s[13], w, N2);
                            if (FP > 0) {
                                const float4* f4 = s_feat + j * (FP / 4);
#pragma unroll
                                for (int v = 0; v < FP / 4; v++) {
                                    const float4 f = f4[v];
                                    E[4 * v + 0] = fma_(f.x, w, E[4 * v + 0]);
                                    E[4 * v + 1] = fma_(f.y, w, E[4 * v + 1]);
                                    E[4 * v + 2] = fma_(f.z, w, E[4 * v + 2]);
                                    E[4 * v + 3] = fma_(f.w, w, E[4 * v + 3]);
                                }
                            }
                            const float4 c = s_rgb[j];
                            C0 = fma_(c.x, w, C0); C1 = fma_(c.y, w, C1); C2 = fma_(c.z, w, C2);
                            T = test_T;
                            last_contributor = contributor;
                        }
                    }
                }
                if (kPairs) {
                    const bool emit = hit && (w >= 0.1f);  // reference: (double)w > 0.1 (forward.cu:422)
                    const unsigned m_emit = __ballot_sync(0xffffffffu, emit);
                    if (m_emit) {
                        const int n_new = __popc(m_emit);
                        if (wcount + n_new > kPairStage) {
                            int gbase = 0;
                            if (lane == 0) gbase = atomicAdd(pair_count, wcount);
                            gbase = __shfl_sync(0xffffffffu, gbase, 0);
                            for (int i = lane; i < wcount; i += 32)
                                if ((int64_t)gbase + i < pair_cap) pairs[gbase + i] = my_pairs[i];
                            __syncwarp();
                            wcount = 0;
                        }
                        if (emit) my_pairs[wcount + __popc(m_emit & ((1u << lane) - 1u))] = make_int2(s_id[j], (int)pix_id);
                        wcount += n_new;
                        __syncwarp();
                    }
                }
            }
        }
    }
    if (kPairs && wcount > 0) {
        int gbase = 0;
        if (lane == 0) gbase = atomicAdd(pair_count, wcount);
        gbase = __shfl_sync(0xffffffffu, gbase, 0);
        for (int i = lane; i < wcount; i += 32)
            if ((int64_t)gbase + i < pair_cap) pairs[gbase + i] = my_pairs[i];
    }

    if (inside) {
        const size_t HW = (size_t)H * W;
        final_T[pix_id] = T;
        final_T[pix_id + HW] = M1;
        final_T[pix_id + 2 * HW] = M2;
        n_contrib[pix_id] = last_contributor;
        n_contrib[pix_id + HW] = median_contributor;
        out_color[pix_id] = fma_(T, __ldg(bg + 0), C0);
        out_color[pix_id + HW] = fma_(T, __ldg(bg + 1), C1);
        out_color[pix_id + 2 * HW] = fma_(T, __ldg(bg + 2), C2);
        out_others[pix_id + 0 * HW] = D;
        out_others[pix_id + 1 * HW] = sub(1.0f, T);
        out_others[pix_id + 2 * HW] = N0;
        out_others[pix_id + 3 * HW] = N1;
        out_others[pix_id + 4 * HW] = N2;
        out_others[pix_id + 5 * HW] = median_depth;
        out_others[pix_id + 6 * HW] = dist;
        if (FP > 0) {
#pragma unroll
            for (int ch = 0; ch < FP; ch++)
                if (ch < F) out_extra[(size_t)ch * HW + pix_id] = E[ch];
        }
    }
}

template <int FP, bool kPairs>
static int launch_one(const IsrForwardArgs& a, cudaStream_t stream) {
    GeomLayout gl(a.P);
    ImageLayout il(a.W, a.H);
    BinLayout bl(a.P, 1, a.W, a.H);  // point_list is at offset 0 regardless of R
    const char* g = static_cast<const char*>(a.geom);
    char* im = static_cast<char*>(a.image);
    const char* b = static_cast<const char*>(a.binning);
    const dim3 grid((a.W + TILE - 1) / TILE, (a.H + TILE - 1) / TILE);
    const size_t smem = FwdSmem<FP>::bytes;
    auto kern = blend_fwd_kernel<FP, kPairs>;
    ISR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, stream>>>(
        reinterpret_cast<const uint2*>(im + il.ranges), reinterpret_cast<const uint32_t*>(b ? b + bl.point_list : nullptr),
        a.W, a.H, a.F, reinterpret_cast<const float4*>(g + gl.splat), reinterpret_cast<const float4*>(g + gl.cull),
        reinterpret_cast<const float4*>(g + gl.rgb), a.extra_attrs, a.background,
        reinterpret_cast<float*>(im + il.final_T), reinterpret_cast<uint32_t*>(im + il.n_contrib), a.out_color,
        a.out_others, a.out_extra, reinterpret_cast<int2*>(a.pairs), a.pair_capacity, a.pair_count);
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

template <bool kPairs>
static int dispatch_F(const IsrForwardArgs& a, cudaStream_t stream) {
    const int F = a.F;
    if (F == 0) return launch_one<0, kPairs>(a, stream);
    if (F <= 4) return launch_one<4, kPairs>(a, stream);
    if (F <= 8) return launch_one<8, kPairs>(a, stream);
    if (F <= 16) return launch_one<16, kPairs>(a, stream);
    if (F <= 24) return launch_one<24, kPairs>(a, stream);
    if (F <= 32) return launch_one<32, kPairs>(a, stream);
    return ISR_ERR_UNSUPPORTED;
}

int launch_blend_fwd(const IsrForwardArgs& a, cudaStream_t stream) {
    const bool want_pairs = a.pairs != nullptr && a.pair_count != nullptr && !(a.flags & ISR_FLAG_NO_PAIRS);
    return want_pairs ? dispatch_F<true>(a, stream) : dispatch_F<false>(a, stream);
}

}  // namespace isr
