// isr_blend_fwd.cu -- K6: per-tile front-to-back alpha compositing of RGB + 7 auxiliary maps + F semantic
// feature channels + the (gaussian, pixel) pair list.   Reference: DSR/cuda_rasterizer/forward.cu:256-462.
//
// Mapping: one thread per pixel, a warp per 8x4 pixel block (8 per 16x16 tile), one warp per CTA.  Warps are
// autonomous (see blend_fwd_kernel): per-warp culling by the per-block footprint bits carried in the list entries
// (computed at emission from a conservative per-Gaussian footprint of K1), cp.async staging of the surviving records (splat 64 B, rgb 16 B, features 4F B) into warp-private shared
// memory, broadcast reads in the inner loop (the reference fetches rgb and features from global per contributing
// (pixel, Gaussian) pair), no block-wide barriers.  Pair-list entries are staged per warp in shared memory and
// flushed with one global atomic per <= 96 pairs (the reference: one global atomic per pair on a single counter).
#include <cstdlib>

#include "isr_common.cuh"

namespace isr {

constexpr int kPairStage = 96;  // per-warp staging slots (int2)

template <int FP>  // feature dim padded to a multiple of 4 (0, 4, 8, 16, 24, 32)
struct FwdSmem {
    static constexpr int kRecF4 = 4 + 1 + FP / 4;  // splat (4 x float4) + rgb (1) + features
    static constexpr size_t per_warp = (size_t)32 * kRecF4 * 16 + 32 * 8 + kPairStage * 8;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
// Shared-memory loads through an explicit 32-bit .shared address (kept in a register by the caller): with generic
// pointers the compiler re-derived the shared window base (S2R SR_CgaCtaId + LEA) inside the inner loop.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int ldsi32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Warp-autonomous blend: every warp owns an 8x4 pixel block and walks its tile's list on its own, 32 entries at a
// time: (1) ids and cull rectangles of the NEXT chunks are prefetched into registers, (2) one lane per entry tests
// the cull rectangle against the warp's pixel block, (3) only the survivors' records are copied (cp.async, 16-byte
// LDGSTS, no registers) into the warp's private shared-memory slots, (4) the survivors are blended in list order
// with broadcast shared-memory reads.  There is no block-wide barrier: warps of a tile neither wait for each other
// nor for the slowest pixel of the tile, and a warp stops as soon as its own 32 pixels are saturated.
// kWarps: warps (8x4 pixel blocks) per CTA; the warps never cooperate, so a CTA is just a scheduling unit.  With
// kWarps = 1 every shared-memory address of the inner loop is `constant + r * record size` (with more warps per CTA
// the compiler re-derived the warp's slot base from S2R/LEA/IMAD in every iteration: 14 of ~165 instructions).
// kWarpsPerSM: occupancy target that sets the register budget (24 -> 80 registers, 28 -> 72, 32 -> 64).
template <int FP, bool kPairs, int kWarpsPerSM, int kWarps, bool kRef>
__global__ void __launch_bounds__(32 * kWarps, kWarpsPerSM / kWarps)
blend_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int F,
                 const float4* __restrict__ splats, const float4* __restrict__ cull4, const float4* __restrict__ cullq,
                 const float4* __restrict__ rgb4, const float* __restrict__ extras, const float* __restrict__ bg,
                 float* __restrict__ final_T,
                 uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_others,
                 float* __restrict__ out_extra, int2* __restrict__ pairs, int64_t pair_cap, int* __restrict__ pair_count,
                 int packed) {
    extern __shared__ __align__(16) float4 smem4[];
    constexpr int REC = FwdSmem<FP>::kRecF4;
    constexpr int kWarpF4 = (int)(FwdSmem<FP>::per_warp / 16);  // float4 units per warp
    static_assert(FwdSmem<FP>::per_warp % 16 == 0, "per-warp shared memory must be a multiple of 16 bytes");
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wic = kWarps == 1 ? 0 : (tid >> 5);            // warp in CTA
    float4* const slots = smem4 + wic * kWarpF4;                            // [32][REC]
    int2* const meta = reinterpret_cast<int2*>(slots + 32 * REC);           // [32] (gaussian id, list index)
    int2* const my_pairs = meta + 32;                                        // [kPairStage]
    // the same two bases as .shared addresses, pinned in registers (the empty asm keeps the compiler from
    // re-deriving them inside the loops)
    uint32_t slots_s = (uint32_t)__cvta_generic_to_shared(slots), meta_s = (uint32_t)__cvta_generic_to_shared(meta);
    asm volatile("" : "+r"(slots_s), "+r"(meta_s));

    const int tiles_x = (W + TILE - 1) / TILE;
    constexpr int kCtasPerTile = 8 / kWarps;
    const int tile_id = blockIdx.x / kCtasPerTile;
    const int warp = (blockIdx.x % kCtasPerTile) * kWarps + wic;  // 8x4 block of the tile: 2 across, 4 down
    const int tile_x = tile_id % tiles_x, tile_y = tile_id / tiles_x;
    const int wx0 = tile_x * TILE + (warp & 1) * 8;
    const int wy0 = tile_y * TILE + (warp >> 1) * 4;
    const int pxi = wx0 + (lane & 7), pyi = wy0 + (lane >> 3);
    const bool inside = pxi < W && pyi < H;
    const uint32_t pix_id = (uint32_t)W * (uint32_t)pyi + (uint32_t)pxi;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const float bx0 = (float)wx0, by0 = (float)wy0;
    const float bx1 = (float)min(wx0 + 7, W - 1), by1 = (float)min(wy0 + 3, H - 1);
    const float bcx = 0.5f * (bx0 + bx1), bcy = 0.5f * (by0 + by1), bhx = 0.5f * (bx1 - bx0), bhy = 0.5f * (by1 - by0);

    const uint2 range = ranges[tile_id];
    const int n_total = (int)(range.y - range.x);
    const uint32_t* __restrict__ plist = point_list + range.x;
    const float c1 = __fdiv_rn(kFar, __fsub_rn(kFar, kNear));

    // `live`: the pixel still blends (inside the image and not saturated).  Kept as a full-width integer: a bool that
    // shares a register with other byte-sized values costs PRMT repacking in every iteration.
    uint32_t live = inside ? 1u : 0u;
    float T = 1.0f;
    // accumulators live in register pairs so that one FFMA2 updates two of them (isr::fma2, bit-identical to two FFMAs)
    float2 C01 = make_float2(0.f, 0.f), C2x = C01, N12 = C01, DM1 = C01, M2dist = C01;  // C2x.y, see below, stays 0
    float N0 = 0, median_depth = 0;
    uint32_t last_contributor = 0, median_contributor = 0;
    float2 E[FP > 0 ? FP / 2 : 1];
#pragma unroll
    for (int ch = 0; ch < FP / 2; ch++) E[ch] = make_float2(0.f, 0.f);
    int wcount = 0;  // staged pairs of this warp (warp-uniform)

    // list entries are fetched one chunk ahead.  Packed entries (isr_common.cuh) carry the "may reach block b of the
    // tile" bits computed at emission; plain entries (>= 2^24 Gaussians) are tested against the footprint data here.
    uint32_t ent_next = (lane < n_total) ? __ldg(plist + lane) : 0u;

    for (int base = 0; base < n_total; base += 32) {
        if (__all_sync(0xffffffffu, live == 0u)) break;
        const uint32_t ent = ent_next;
        const bool have = base + lane < n_total;
        ent_next = (base + 32 + lane < n_total) ? __ldg(plist + base + 32 + lane) : 0u;
        int id;
        bool ov;
        if (packed) {
            id = (int)(ent & kIdMask);
            ov = have && ((ent >> (kIdBits + warp)) & 1u);
        } else {
            id = (int)ent;
            ov = have;
            if (ov) {
                const float4 cr = __ldg(cull4 + id);
                ov = !(cr.z < bx0 || cr.x > bx1 || cr.w < by0 || cr.y > by1);
            }
            if (ov) {  // second stage: the conic itself against the block (corner overlaps of the rectangle)
                const float4* q = cullq + (size_t)id * 3;
                ov = !block_outside(__ldg(q), __ldg(q + 1), __ldg(q + 2).x, bcx, bcy, bhx, bhy);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ov);
        if (m == 0) continue;
        if (ov) {
            const int rank = __popc(m & ((1u << lane) - 1u));
            float4* dst = slots + rank * REC;
            const float4* sp = splats + (size_t)id * 4;
            cp_async16(dst + 0, sp + 0);
            cp_async16(dst + 1, sp + 1);
            cp_async16(dst + 2, sp + 2);
            cp_async16(dst + 3, sp + 3);
            cp_async16(dst + 4, rgb4 + id);
            if (FP > 0) {
                if ((F & 3) == 0) {
                    const float4* fp = reinterpret_cast<const float4*>(extras + (size_t)id * F);
#pragma unroll
                    for (int v = 0; v < FP / 4; v++) {
                        if (v * 4 < F) cp_async16(dst + 5 + v, fp + v);
                        else dst[5 + v] = make_float4(0, 0, 0, 0);
                    }
                } else {
                    float* dstf = reinterpret_cast<float*>(dst + 5);
#pragma unroll
                    for (int ch = 0; ch < FP; ch++) dstf[ch] = (ch < F) ? __ldg(extras + (size_t)id * F + ch) : 0.0f;
                }
            }
            meta[rank] = make_int2(id, base + lane + 1);  // .y = the reference's 1-based `contributor` of this entry
        }
        cp_async_wait_all();
        __syncwarp();
        const int n_surv = __popc(m);
        uint32_t rec_s = slots_s;  // .shared address of survivor r's record
        for (int r = 0; r < n_surv; r++, rec_s += REC * 16) {
            float w = 0.0f;  // blend weight of this (pixel, Gaussian) pair; stays 0 if the pair does not contribute
            if (live) {
                float s[16];
                *reinterpret_cast<float4*>(s + 0) = lds128(rec_s);
                *reinterpret_cast<float4*>(s + 4) = lds128(rec_s + 16);
                *reinterpret_cast<float4*>(s + 8) = lds128(rec_s + 32);
                *reinterpret_cast<float2*>(s + 14) = lds64(rec_s + 56);  // opacity, power_cut
                PairEval e;
                if (eval_pair<kRef, false>(pixx, pixy, s, e)) {
                    const float test_T = mul(T, sub(1.0f, e.alpha));
                    if (test_T < kTMin) {
                        live = 0u;
                    } else {
                        const uint32_t contributor = (uint32_t)ldsi32(meta_s + 8 * r + 4);
                        w = mul(e.alpha, T);
                        const float A = sub(1.0f, T);
                        // [sass] m = (1 + (-near)/depth) * (far/(far-near)): div.rn, FADD, FMUL
                        // (depth >= near here; the fast path of the division is exact up to 2^60)
                        const float nd = e.depth <= 0x1p60f ? div_fast(kNear, e.depth) : __fdiv_rn(kNear, e.depth);
                        const float mdep = mul(c1, sub(1.0f, nd));
                        const float mm = mul(mdep, mdep);
                        const float dt = fma_(-add(mdep, mdep), DM1.y, fma_(mm, A, M2dist.x));
                        M2dist = fma2(make_float2(mm, dt), w, M2dist);       // M2 += mm*w, dist += dt*w
                        DM1 = fma2(make_float2(e.depth, mdep), w, DM1);      // D += depth*w, M1 += mdep*w
                        if (T > 0.5f) { median_depth = e.depth; median_contributor = contributor; }
                        N0 = fma_(s[11], w, N0);
                        N12 = fma2(lds64(rec_s + 48), w, N12);
                        if (FP > 0) {
                            // [sass] E[ch] = fma(T, alpha * feature, E[ch])  (forward.cu:415: extras * alpha * T)
#pragma unroll
                            for (int v = 0; v < FP / 4; v++) {
                                const float4 f = lds128(rec_s + 80 + 16 * v);
                                E[2 * v + 0] = fma2(mul2(make_float2(f.x, f.y), e.alpha), T, E[2 * v + 0]);
                                E[2 * v + 1] = fma2(mul2(make_float2(f.z, f.w), e.alpha), T, E[2 * v + 1]);
                            }
                        }
                        const float4 c = lds128(rec_s + 64);  // (r, g, b, 0)
                        C01 = fma2(make_float2(c.x, c.y), w, C01);
                        C2x = fma2(make_float2(c.z, c.w), w, C2x);
                        T = test_T;
                        last_contributor = contributor;
                    }
                }
            }
            if (kPairs) {
                // reference: (double)w > 0.1 (forward.cu:422)  <=>  w >= 0.1f
                const unsigned m_emit = __ballot_sync(0xffffffffu, w >= 0.1f);
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
                    if (w >= 0.1f) my_pairs[wcount + __popc(m_emit & ((1u << lane) - 1u))] = make_int2(ldsi32(meta_s + 8 * r), (int)pix_id);
                    wcount += n_new;
                    __syncwarp();
                }
            }
        }
        __syncwarp();  // all lanes are done reading the slots before the next chunk overwrites them
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
        final_T[pix_id + HW] = DM1.y;
        final_T[pix_id + 2 * HW] = M2dist.x;
        n_contrib[pix_id] = last_contributor;
        n_contrib[pix_id + HW] = median_contributor;
        out_color[pix_id] = fma_(T, __ldg(bg + 0), C01.x);
        out_color[pix_id + HW] = fma_(T, __ldg(bg + 1), C01.y);
        out_color[pix_id + 2 * HW] = fma_(T, __ldg(bg + 2), C2x.x);
        out_others[pix_id + 0 * HW] = DM1.x;
        out_others[pix_id + 1 * HW] = sub(1.0f, T);
        out_others[pix_id + 2 * HW] = N0;
        out_others[pix_id + 3 * HW] = N12.x;
        out_others[pix_id + 4 * HW] = N12.y;
        out_others[pix_id + 5 * HW] = median_depth;
        out_others[pix_id + 6 * HW] = M2dist.y;
        if (FP > 0) {
#pragma unroll
            for (int ch = 0; ch < FP; ch++)
                if (ch < F) out_extra[(size_t)ch * HW + pix_id] = (ch & 1) ? E[ch >> 1].y : E[ch >> 1].x;
        }
    }
}

template <int FP, bool kPairs>
static int launch_one(const IsrForwardArgs& a, cudaStream_t stream) {
    GeomLayout gl(a.P);
    ImageLayout il(a.W, a.H);
    const char* g = static_cast<const char*>(a.geom);
    char* im = static_cast<char*>(a.image);
    const char* b = static_cast<const char*>(a.binning);
    const int num_tiles = ((a.W + TILE - 1) / TILE) * ((a.H + TILE - 1) / TILE);
    // Occupancy = register budget: 32 warps/SM (64 registers), 28 (72) or 24 (80); the largest that compiles without
    // spills for this F (measured at cfg3, F=16: 28 beats 24 by 5% and 32 by 6%).
    constexpr int kWps = FP <= 8 ? 32 : (FP <= 24 ? 28 : 24);
    auto launch = [&](auto kern, int w) -> int {
        const size_t smem = (size_t)w * FwdSmem<FP>::per_warp;
        ISR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<num_tiles * (8 / w), 32 * w, smem, stream>>>(
            reinterpret_cast<const uint2*>(im + il.ranges), reinterpret_cast<const uint32_t*>(b),  // point_list sits at offset 0 of the binning workspace
            a.W, a.H, a.F, reinterpret_cast<const float4*>(g + gl.splat), reinterpret_cast<const float4*>(g + gl.cull),
            reinterpret_cast<const float4*>(g + gl.cullq), reinterpret_cast<const float4*>(g + gl.rgb), a.extra_attrs, a.background,
            reinterpret_cast<float*>(im + il.final_T), reinterpret_cast<uint32_t*>(im + il.n_contrib), a.out_color,
            a.out_others, a.out_extra, reinterpret_cast<int2*>(a.pairs), a.pair_capacity, a.pair_count,
            entries_packed(a.P) ? 1 : 0); note_launch();
        ISR_CUDA_TRY(cudaGetLastError());
        return ISR_OK;
    };
    if (a.flags & ISR_FLAG_SPEC_ARITH) return launch(blend_fwd_kernel<FP, kPairs, kWps, 1, false>, 1);
    static const int env_w = [] { const char* e = getenv("ISR_FWD_WARPS"); return e ? atoi(e) : 1; }();  // experiment switch
    if (env_w == 2) return launch(blend_fwd_kernel<FP, kPairs, kWps, 2, true>, 2);
    return launch(blend_fwd_kernel<FP, kPairs, kWps, 1, true>, 1);
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
