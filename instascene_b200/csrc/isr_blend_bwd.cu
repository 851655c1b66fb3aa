// isr_blend_bwd.cu -- K7: backward of the per-tile blend.  Reference: DSR/cuda_rasterizer/backward.cu:143-466.
//
// Dense kernel (all gradients, dense CHW cotangents): one CTA per 16x16 tile, one thread per pixel, a warp per
// 8x4 pixel block, list replayed back-to-front from the per-pixel state saved by the forward.  Where the
// reference issues 21+F global float atomics per contributing (pixel, Gaussian) pair -- 32 lanes hitting the
// same address -- this kernel first reduces the 18+F per-Gaussian partial gradients across the warp with a
// transposing butterfly (31 shuffles for 32 values: after the last step lane L owns the total of value L) and
// then issues ONE red.global per value per (warp, Gaussian), skipping Gaussians no lane of the warp touched.
//
// Sparse kernel (isr_backward_extra_sparse): when only the semantic feature is trainable (train_semantic.py)
// the only non-zero cotangent is dL/d(extra map) at <= 32768 sampled pixels, and
//   dL/d extra[g][ch] = sum_pixels w(g,pix) * dL/dE[ch](pix)            (backward.cu:401)
// needs just the forward weights w = alpha*T, so one warp per sampled pixel walks that pixel's tile list
// front-to-back with the lanes spread over 32 consecutive Gaussians and a warp-shuffle product scan for T.
#include "isr_common.cuh"

namespace isr {

template <int N, int OFF>
struct Butterfly {
    __device__ __forceinline__ static void run(float* v, int lane) {
        if constexpr (N > 1) {
            constexpr int h = N / 2;
            const bool up = (lane & OFF) != 0;
#pragma unroll
            for (int i = 0; i < h; i++) {
                const float send = up ? v[i] : v[i + h];
                const float keep = up ? v[i + h] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
            }
            if constexpr (OFF > 1) Butterfly<h, OFF / 2>::run(v, lane);
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], OFF);
            if constexpr (OFF > 1) Butterfly<1, OFF / 2>::run(v, lane);
        }
    }
};
// After warp_transpose_reduce<N>(v): v[0] on lane L is the warp-wide sum of value index (L >> log2(32/N)).
template <int N>
__device__ __forceinline__ float warp_transpose_reduce(float* v, int lane) {
    Butterfly<N, 16>::run(v, lane);
    return v[0];
}

template <int FP>
struct BwdSmem {
    static constexpr int kRecF4 = 4 + 1 + FP / 4;  // splat (4 x float4) + rgb (1) + features
    static constexpr size_t per_warp = (size_t)32 * kRecF4 * 16 + 32 * 8;
};

__device__ __forceinline__ void cp_async16_b(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}

// Same warp-autonomous structure as blend_fwd_kernel, walking the list BACK to front: per-warp cull-rectangle test
// (exact: a culled Gaussian contributed to no pixel of the block in the forward), cp.async staging of survivors.
// kWarps: warps (8x4 pixel blocks) per CTA, as in blend_fwd_kernel (the warps never cooperate).
template <int FP, int kWarps, bool kRef>
__global__ void __launch_bounds__(32 * kWarps, (FP == 0 ? 24 : 8) / kWarps)
blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int F,
                 const float* __restrict__ bg, const float4* __restrict__ splats, const float4* __restrict__ cull4,
                 const float4* __restrict__ cullq, const float4* __restrict__ rgb4, const float* __restrict__ extras,
                 const float* __restrict__ final_Ts,
                 const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
                 const float* __restrict__ dL_dothers, const float* __restrict__ dL_dpix_extra,
                 float* __restrict__ dL_dtransMat, float* __restrict__ dL_dmean2D, float* __restrict__ dL_dnormal3D,
                 float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolors, float* __restrict__ dL_dextras,
                 int packed) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int REC = BwdSmem<FP>::kRecF4;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    unsigned char* wbase = smem_raw + (size_t)(tid >> 5) * BwdSmem<FP>::per_warp;
    float4* slots = reinterpret_cast<float4*>(wbase);                     // [32][REC]
    int2* meta = reinterpret_cast<int2*>(wbase + (size_t)32 * REC * 16);  // [32] (gaussian id, list index)

    const int tiles_x = (W + TILE - 1) / TILE;
    constexpr int kCtasPerTile = 8 / kWarps;
    const int tile_id = blockIdx.x / kCtasPerTile;
    const int warp = (blockIdx.x % kCtasPerTile) * kWarps + (tid >> 5);  // 8x4 block of the tile: 2 across, 4 down
    const int wx0 = (tile_id % tiles_x) * TILE + (warp & 1) * 8;
    const int wy0 = (tile_id / tiles_x) * TILE + (warp >> 1) * 4;
    const int pxi = wx0 + (lane & 7), pyi = wy0 + (lane >> 3);
    const bool inside = pxi < W && pyi < H;
    const size_t HW = (size_t)H * W;
    const size_t pix_id = (size_t)W * pyi + pxi;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const float bx0 = (float)wx0, by0 = (float)wy0;
    const float bx1 = (float)min(wx0 + 7, W - 1), by1 = (float)min(wy0 + 3, H - 1);
    const float bcx = 0.5f * (bx0 + bx1), bcy = 0.5f * (by0 + by1), bhx = 0.5f * (bx1 - bx0), bhy = 0.5f * (by1 - by0);

    const uint2 range = ranges[tile_id];
    const int n_total = (int)(range.y - range.x);
    const uint32_t* __restrict__ plist = point_list + range.x;
    const float c1 = __fdiv_rn(kFar, __fsub_rn(kFar, kNear));
    const float c3 = __fdiv_rn(__fmul_rn(kFar, kNear), __fsub_rn(kFar, kNear));

    const float T_final = inside ? final_Ts[pix_id] : 0.0f;
    float T = T_final;
    const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0u;
    const uint32_t median_contributor = inside ? n_contrib[pix_id + HW] : 0u;
    float dLdC0 = 0, dLdC1 = 0, dLdC2 = 0, dL_ddepth = 0, dL_daccum = 0, dL_dreg = 0, dL_dmedian = 0;
    float dLdN0 = 0, dLdN1 = 0, dLdN2 = 0;
    float dLdE[FP > 0 ? FP : 1];
#pragma unroll
    for (int ch = 0; ch < FP; ch++) dLdE[ch] = 0.0f;
    if (inside) {
        if (dL_dpixels) { dLdC0 = dL_dpixels[pix_id]; dLdC1 = dL_dpixels[pix_id + HW]; dLdC2 = dL_dpixels[pix_id + 2 * HW]; }
        if (dL_dothers) {
            dL_ddepth = dL_dothers[pix_id];
            dL_daccum = dL_dothers[pix_id + HW];
            dLdN0 = dL_dothers[pix_id + 2 * HW]; dLdN1 = dL_dothers[pix_id + 3 * HW]; dLdN2 = dL_dothers[pix_id + 4 * HW];
            dL_dmedian = dL_dothers[pix_id + 5 * HW];
            dL_dreg = dL_dothers[pix_id + 6 * HW];
        }
        if (FP > 0 && dL_dpix_extra) {
#pragma unroll
            for (int ch = 0; ch < FP; ch++) if (ch < F) dLdE[ch] = dL_dpix_extra[(size_t)ch * HW + pix_id];
        }
    }
    const float final_D = inside ? final_Ts[pix_id + HW] : 0.0f;
    const float final_D2 = inside ? final_Ts[pix_id + 2 * HW] : 0.0f;
    const float final_A = 1.0f - T_final;
    const float bg_dot = fma_(__ldg(bg + 2), dLdC2, fma_(__ldg(bg + 1), dLdC1, mul(__ldg(bg + 0), dLdC0)));

    float acc_c0 = 0, acc_c1 = 0, acc_c2 = 0, last_c0 = 0, last_c1 = 0, last_c2 = 0;
    float acc_depth = 0, last_depth = 0, acc_alpha = 0;
    float acc_n0 = 0, acc_n1 = 0, acc_n2 = 0, last_n0 = 0, last_n1 = 0, last_n2 = 0;
    float acc_e[FP > 0 ? FP : 1], last_e[FP > 0 ? FP : 1];
#pragma unroll
    for (int ch = 0; ch < FP; ch++) { acc_e[ch] = 0.0f; last_e[ch] = 0.0f; }
    float last_dL_dT = 0, last_alpha = 0;

    // Destination of the butterfly result held by this lane (loop invariant): lanes 2i and 2i+1 end up with the total
    // of value i = lane>>1: 0-2 colour, 3-5 normal, 6-14 transMat, 15 opacity.
    float* lane_dst = nullptr;
    int lane_stride = 0;
    {
        const int vi = lane >> 1;
        if ((lane & 1) == 0) {
            if (vi < 3) { lane_dst = dL_dcolors ? dL_dcolors + vi : nullptr; lane_stride = 3; }
            else if (vi < 6) { lane_dst = dL_dnormal3D ? dL_dnormal3D + (vi - 3) : nullptr; lane_stride = 3; }
            else if (vi < 15) { lane_dst = dL_dtransMat ? dL_dtransMat + (vi - 6) : nullptr; lane_stride = 9; }
            else { lane_dst = dL_dopacity; lane_stride = 1; }
        }
    }

    // the highest list index any pixel of this WARP needs
    int n_need = (int)last_contributor;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) n_need = max(n_need, __shfl_xor_sync(0xffffffffu, n_need, off));
    n_need = min(n_need, n_total);

    // chunk with top index `top` covers list indices top-1-lane (lane 0 = last entry); entries one chunk ahead
    auto idx_of = [&](int top) { return top - 1 - lane; };
    uint32_t ent_next = (idx_of(n_need) >= 0) ? __ldg(plist + idx_of(n_need)) : 0u;

    for (int top = n_need; top > 0; top -= 32) {
        const uint32_t ent = ent_next;
        const bool have = idx_of(top) >= 0;
        ent_next = (idx_of(top - 32) >= 0) ? __ldg(plist + idx_of(top - 32)) : 0u;
        int id;
        bool ov;
        if (packed) {  // per-block footprint bits computed at emission (isr_common.cuh)
            id = (int)(ent & kIdMask);
            ov = have && ((ent >> (kIdBits + warp)) & 1u);
        } else {
            id = (int)ent;
            ov = have;
            if (ov) {
                const float4 cr = __ldg(cull4 + id);
                ov = !(cr.z < bx0 || cr.x > bx1 || cr.w < by0 || cr.y > by1);
            }
            if (ov) {
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
            cp_async16_b(dst + 0, sp + 0);
            cp_async16_b(dst + 1, sp + 1);
            cp_async16_b(dst + 2, sp + 2);
            cp_async16_b(dst + 3, sp + 3);
            cp_async16_b(dst + 4, rgb4 + id);
            if (FP > 0) {
                if ((F & 3) == 0) {
                    const float4* fp = reinterpret_cast<const float4*>(extras + (size_t)id * F);
#pragma unroll
                    for (int v = 0; v < FP / 4; v++) {
                        if (v * 4 < F) cp_async16_b(dst + 5 + v, fp + v);
                        else dst[5 + v] = make_float4(0, 0, 0, 0);
                    }
                } else {
                    float* dstf = reinterpret_cast<float*>(dst + 5);
#pragma unroll
                    for (int ch = 0; ch < FP; ch++) dstf[ch] = (ch < F) ? __ldg(extras + (size_t)id * F + ch) : 0.0f;
                }
            }
            meta[rank] = make_int2(id, idx_of(top));
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        __syncwarp();
        const int n_surv = __popc(m);

        for (int t = 0; t < n_surv; t++) {
            const int2 mt = meta[t];
            const uint32_t contributor = (uint32_t)mt.y;  // 0-based list index
            const bool active = contributor < last_contributor;
            const float* s = reinterpret_cast<const float*>(slots + t * REC);
            PairEval e;
            const bool hit = active && eval_pair<kRef, true>(pixx, pixy, s, e);
            if (!__any_sync(0xffffffffu, hit)) continue;
            // per-lane partial gradients (0 when this lane does not contribute)
            // v[0..2] colour, v[3..5] normal, v[6..14] transMat, v[15] opacity; mean2D (2D-filter branch only) apart
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = 0.0f;
            float m2x = 0.0f, m2y = 0.0f;
            float ve[FP > 0 ? FP : 1];
#pragma unroll
            for (int ch = 0; ch < FP; ch++) ve[ch] = 0.0f;
            if (hit) {
                const float alpha = e.alpha, G = e.G, c_d = e.depth;
                const float ra = rcp_fast(sub(1.0f, alpha));
                T = mul(T, ra);
                const float w = mul(alpha, T);
                float dL_dalpha = 0.0f;
                const float one_m_la = sub(1.0f, last_alpha);
                const float4 col = slots[t * REC + 4];
                acc_c0 = fma_(last_alpha, last_c0, mul(one_m_la, acc_c0)); last_c0 = col.x;
                dL_dalpha = fma_(sub(col.x, acc_c0), dLdC0, dL_dalpha);
                acc_c1 = fma_(last_alpha, last_c1, mul(one_m_la, acc_c1)); last_c1 = col.y;
                dL_dalpha = fma_(sub(col.y, acc_c1), dLdC1, dL_dalpha);
                acc_c2 = fma_(last_alpha, last_c2, mul(one_m_la, acc_c2)); last_c2 = col.z;
                dL_dalpha = fma_(sub(col.z, acc_c2), dLdC2, dL_dalpha);
                v[0] = mul(w, dLdC0); v[1] = mul(w, dLdC1); v[2] = mul(w, dLdC2);

                float dL_dz = 0.0f;
                const float rcd = rcp_fast(c_d);
                const float m_d = mul(c1, sub(1.0f, mul(kNear, rcd)));
                const float dmd_dd = mul(mul(c3, rcd), rcd);
                if (contributor == median_contributor - 1u) dL_dz = add(dL_dz, dL_dmedian);
                const float mm = mul(m_d, m_d);
                const float dL_dweight = mul(fma_(-add(m_d, m_d), final_D, fma_(mm, final_A, final_D2)), dL_dreg);
                dL_dalpha = add(dL_dalpha, sub(dL_dweight, last_dL_dT));
                last_dL_dT = fma_(dL_dweight, alpha, mul(sub(1.0f, alpha), last_dL_dT));
                const float dL_dmd = mul(mul(add(w, w), fma_(m_d, final_A, -final_D)), dL_dreg);
                dL_dz = fma_(dL_dmd, dmd_dd, dL_dz);
                acc_depth = fma_(last_alpha, last_depth, mul(one_m_la, acc_depth));
                last_depth = c_d;
                dL_dalpha = fma_(sub(c_d, acc_depth), dL_ddepth, dL_dalpha);
                acc_alpha = add(last_alpha, mul(one_m_la, acc_alpha));
                dL_dalpha = fma_(sub(1.0f, acc_alpha), dL_daccum, dL_dalpha);
                acc_n0 = fma_(last_alpha, last_n0, mul(one_m_la, acc_n0)); last_n0 = s[11];
                dL_dalpha = fma_(sub(s[11], acc_n0), dLdN0, dL_dalpha);
                acc_n1 = fma_(last_alpha, last_n1, mul(one_m_la, acc_n1)); last_n1 = s[12];
                dL_dalpha = fma_(sub(s[12], acc_n1), dLdN1, dL_dalpha);
                acc_n2 = fma_(last_alpha, last_n2, mul(one_m_la, acc_n2)); last_n2 = s[13];
                dL_dalpha = fma_(sub(s[13], acc_n2), dLdN2, dL_dalpha);
                v[3] = mul(w, dLdN0); v[4] = mul(w, dLdN1); v[5] = mul(w, dLdN2);
                if (FP > 0) {
                    const float* f = reinterpret_cast<const float*>(slots + t * REC + 5);
#pragma unroll
                    for (int ch = 0; ch < FP; ch++) {
                        const float ex = f[ch];
                        acc_e[ch] = fma_(last_alpha, last_e[ch], mul(one_m_la, acc_e[ch]));
                        last_e[ch] = ex;
                        dL_dalpha = fma_(sub(ex, acc_e[ch]), dLdE[ch], dL_dalpha);
                        ve[ch] = mul(w, dLdE[ch]);
                    }
                }
                dL_dalpha = mul(dL_dalpha, T);
                last_alpha = alpha;
                dL_dalpha = fma_(mul(-T_final, ra), bg_dot, dL_dalpha);
                const float dL_dG = mul(s[14], dL_dalpha);
                dL_dz = fma_(w, dL_ddepth, dL_dz);
                if (e.use3d) {
                    const float nGd = mul(dL_dG, -G);
                    const float dsx = fma_(nGd, e.sx, mul(dL_dz, s[6]));
                    const float dsy = fma_(nGd, e.sy, mul(dL_dz, s[7]));
                    const float rpz = __frcp_rn(e.pz);
                    const float dpx = mul(dsx, rpz), dpy = mul(dsy, rpz);
                    const float dpz = -fma_(dpx, e.sx, mul(dpy, e.sy));
                    const float dkx = fma_(e.ly, dpz, -mul(e.lz, dpy));
                    const float dky = fma_(e.lz, dpx, -mul(e.lx, dpz));
                    const float dkz = fma_(e.lx, dpy, -mul(e.ly, dpx));
                    const float dlx = fma_(dpy, e.kz, -mul(dpz, e.ky));
                    const float dly = fma_(dpz, e.kx, -mul(dpx, e.kz));
                    const float dlz = fma_(dpx, e.ky, -mul(dpy, e.kx));
                    v[6] = -dkx; v[7] = -dky; v[8] = -dkz;
                    v[9] = -dlx; v[10] = -dly; v[11] = -dlz;
                    v[12] = fma_(dL_dz, e.sx, fma_(pixx, dkx, mul(pixy, dlx)));
                    v[13] = fma_(dL_dz, e.sy, fma_(pixx, dky, mul(pixy, dly)));
                    v[14] = add(fma_(pixx, dkz, mul(pixy, dlz)), dL_dz);
                } else {
                    const float nG2 = mul(-G, kFilterInvSquare);
                    m2x = mul(dL_dG, mul(nG2, e.ddx));
                    m2y = mul(dL_dG, mul(nG2, e.ddy));
                    v[14] = dL_dz;
                }
                v[15] = mul(G, dL_dalpha);
            }
            const int g = mt.x;
            // 16-value transposing butterfly: afterwards lanes 2i and 2i+1 both hold the total of value i
            const float tot = warp_transpose_reduce<16>(v, lane);
            if (lane_dst != nullptr && tot != 0.0f) atomicAdd(lane_dst + (size_t)g * lane_stride, tot);
            // the low-pass (2D) branch is rare: reduce its two values only when some lane took it
            if (__any_sync(0xffffffffu, hit && !e.use3d)) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    m2x += __shfl_xor_sync(0xffffffffu, m2x, off);
                    m2y += __shfl_xor_sync(0xffffffffu, m2y, off);
                }
                if (lane == 0 && dL_dmean2D) {
                    if (m2x != 0.0f) atomicAdd(dL_dmean2D + (size_t)g * 3 + 0, m2x);
                    if (m2y != 0.0f) atomicAdd(dL_dmean2D + (size_t)g * 3 + 1, m2y);
                }
            }
            if (FP > 0) {
                constexpr int NE = FP <= 8 ? 8 : (FP <= 16 ? 16 : 32);
                float vp[NE];
#pragma unroll
                for (int i = 0; i < NE; i++) vp[i] = (i < FP) ? ve[i] : 0.0f;
                const float tote = warp_transpose_reduce<NE>(vp, lane);
                constexpr int shift = NE == 32 ? 0 : (NE == 16 ? 1 : 2);
                const int ch = lane >> shift;
                if ((lane & ((1 << shift) - 1)) == 0 && ch < F && tote != 0.0f && dL_dextras)
                    atomicAdd(dL_dextras + (size_t)g * F + ch, tote);
            }
        }
        __syncwarp();  // slots are reused by the next chunk
    }
}

// ---------------------------------------------------------------------------------------------------------
// Sampled pixels only (a warp per sample; up to ISR_MAX_SPARSE_VIEWS prepared views in one launch).
// ---------------------------------------------------------------------------------------------------------
// Forward: the features of one pixel, composited front to back with the arithmetic and the order of blend_fwd_kernel
// (forward.cu:355-437 restricted to the extra channels): bit-identical to the dense map at that pixel.  Per chunk of 32
// list entries the lanes evaluate one entry each (alpha), the contributing lanes park their Gaussian's F features in the
// warp's shared-memory slots, then the contributors are blended serially in list order with lane c holding channel c --
// the transmittance chain and every accumulation keep the dense kernel's sequence of roundings.  Stops at saturation
// and writes n_contrib (1-based index of the last contributor) and the final transmittance of the pixel.
template <int FP, bool kRef>
__global__ void __launch_bounds__(256)
extra_sparse_fwd_kernel(const SparseViewsDev v, int n, const int* __restrict__ pix_ids, const int* __restrict__ view_ids,
                        int W, int H, int F, const float* __restrict__ extras, float* __restrict__ out, int packed) {
    extern __shared__ __align__(16) float s_feat[];  // [8 warps][32 entries][FP]
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    if (warp_global >= n) return;
    float* slots = s_feat + (size_t)wic * 32 * FP;
    const int pix = pix_ids[warp_global];
    const int vw = view_ids ? view_ids[warp_global] : 0;
    if (pix < 0 || pix >= W * H || vw < 0 || vw >= v.n_views) {
        if (lane < F) out[(size_t)warp_global * F + lane] = 0.0f;
        return;
    }
    const int pxi = pix % W, pyi = pix / W;
    const int tiles_x = (W + TILE - 1) / TILE;
    const uint2 range = v.ranges[vw][(pyi / TILE) * tiles_x + (pxi / TILE)];
    const int n_total = (int)(range.y - range.x);
    const float4* __restrict__ splats = v.splats[vw];
    const float4* __restrict__ cull4 = v.cull4[vw];
    const uint32_t* __restrict__ plist = v.point_list[vw] + range.x;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const int blk = ((pyi % TILE) / 4) * 2 + (pxi % TILE) / 8;  // the pixel's 8x4 block of its tile
    float T = 1.0f, E = 0.0f;  // lane c accumulates channel c (warp-uniform T)
    uint32_t last = 0;
    bool done = false;
    uint32_t ent_next = (lane < n_total) ? __ldg(plist + lane) : 0u;  // entries one chunk ahead
    for (int base = 0; base < n_total && !done; base += 32) {
        const uint32_t ent = ent_next;
        const bool have = base + lane < n_total;
        ent_next = (base + 32 + lane < n_total) ? __ldg(plist + base + 32 + lane) : 0u;
        float alpha = 0.0f;
        int g;
        bool cand;
        if (packed) {  // the entry says whether the Gaussian can reach this pixel's block at all
            g = (int)(ent & kIdMask);
            cand = have && ((ent >> (kIdBits + blk)) & 1u);
        } else {  // exact pre-test: the pixel lies outside the Gaussian's conservative cull rectangle
            g = (int)ent;
            cand = have;
            if (cand) {
                const float4 cr = __ldg(cull4 + g);
                cand = !(cr.z < pixx || cr.x > pixx || cr.w < pixy || cr.y > pixy);
            }
        }
        if (!__any_sync(0xffffffffu, cand)) continue;
        if (cand) {
            float s[16];
            const float4* sp = splats + (size_t)g * 4;
            *reinterpret_cast<float4*>(s + 0) = __ldg(sp + 0);
            *reinterpret_cast<float4*>(s + 4) = __ldg(sp + 1);
            *reinterpret_cast<float4*>(s + 8) = __ldg(sp + 2);
            *reinterpret_cast<float4*>(s + 12) = __ldg(sp + 3);
            PairEval e;
            if (eval_pair<kRef, false>(pixx, pixy, s, e)) alpha = e.alpha;
        }
        unsigned contrib = __ballot_sync(0xffffffffu, alpha != 0.0f);
        if (contrib == 0u) continue;
        if (alpha != 0.0f) {  // park this Gaussian's features where every lane can read its channel
            float* dst = slots + lane * FP;
            if ((F & 3) == 0) {
                const float4* src = reinterpret_cast<const float4*>(extras + (size_t)g * F);
#pragma unroll
                for (int q = 0; q < FP / 4; q++)
                    if (q * 4 < F) *reinterpret_cast<float4*>(dst + 4 * q) = __ldg(src + q);
            } else {
#pragma unroll
                for (int ch = 0; ch < FP; ch++)
                    if (ch < F) dst[ch] = __ldg(extras + (size_t)g * F + ch);
            }
        }
        __syncwarp();
        while (contrib) {  // contributing entries only, in list order
            const int l = __ffs(contrib) - 1;
            contrib &= contrib - 1;
            const float a_l = __shfl_sync(0xffffffffu, alpha, l);
            const float test_T = mul(T, sub(1.0f, a_l));
            if (test_T < kTMin) {  // forward.cu:395-399: this entry is not blended and the pixel is finished
                done = true;
                break;
            }
            const float f = (lane < F) ? slots[l * FP + (lane < FP ? lane : 0)] : 0.0f;
            E = fma_(mul(f, a_l), T, E);  // forward.cu:415 as compiled: fma(feature * alpha, T, E)
            T = test_T;
            last = (uint32_t)(base + l + 1);
        }
        __syncwarp();  // the slots are rewritten by the next chunk
    }
    if (lane < F) out[(size_t)warp_global * F + lane] = E;
    if (lane == 0) {
        v.n_contrib[vw][pix] = last;
        v.final_T[vw][pix] = T;
    }
}

// Backward of the same: dL/d(extra_attrs) from the cotangent rows of the samples.
template <int FP, bool kRef>
__global__ void __launch_bounds__(256)
extra_sparse_bwd_kernel(const SparseViewsDev v, int n, const int* __restrict__ pix_ids, const int* __restrict__ view_ids,
                        const float* __restrict__ dLdE_samples, int W, int H, int F, float* __restrict__ dL_dextras,
                        int packed) {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_global >= n) return;
    const int pix = pix_ids[warp_global];
    const int vw = view_ids ? view_ids[warp_global] : 0;
    if (pix < 0 || pix >= W * H || vw < 0 || vw >= v.n_views) return;
    const int pxi = pix % W, pyi = pix / W;
    const int tiles_x = (W + TILE - 1) / TILE;
    const uint2 range = v.ranges[vw][(pyi / TILE) * tiles_x + (pxi / TILE)];
    const float4* __restrict__ splats = v.splats[vw];
    const float4* __restrict__ cull4 = v.cull4[vw];
    const int last = (int)v.n_contrib[vw][pix];  // number of list entries up to and including the last contributor
    const float pixx = (float)pxi, pixy = (float)pyi;
    float dE[FP];
#pragma unroll
    for (int ch = 0; ch < FP; ch++) dE[ch] = (ch < F) ? dLdE_samples[(size_t)warp_global * F + ch] : 0.0f;
    float T = 1.0f;  // transmittance in front of the current group of 32 (warp-uniform)
    const uint32_t* __restrict__ plist = v.point_list[vw] + range.x;
    const int blk = ((pyi % TILE) / 4) * 2 + (pxi % TILE) / 8;  // the pixel's 8x4 block of its tile
    uint32_t ent_next = (lane < last) ? __ldg(plist + lane) : 0u;  // entries one chunk ahead
    for (int base = 0; base < last; base += 32) {
        const uint32_t ent = ent_next;
        const bool have = base + lane < last;
        ent_next = (base + 32 + lane < last) ? __ldg(plist + base + 32 + lane) : 0u;
        float alpha = 0.0f;
        int g;
        bool cand;
        if (packed) {  // the entry says whether the Gaussian can reach this pixel's block at all
            g = (int)(ent & kIdMask);
            cand = have && ((ent >> (kIdBits + blk)) & 1u);
        } else {  // exact pre-test: the pixel lies outside the Gaussian's conservative cull rectangle
            g = (int)ent;
            cand = have;
            if (cand) {
                const float4 cr = __ldg(cull4 + g);
                cand = !(cr.z < pixx || cr.x > pixx || cr.w < pixy || cr.y > pixy);
            }
        }
        if (!__any_sync(0xffffffffu, cand)) continue;
        if (cand) {
            float s[16];
            const float4* sp = splats + (size_t)g * 4;
            *reinterpret_cast<float4*>(s + 0) = __ldg(sp + 0);
            *reinterpret_cast<float4*>(s + 4) = __ldg(sp + 1);
            *reinterpret_cast<float4*>(s + 8) = __ldg(sp + 2);
            *reinterpret_cast<float4*>(s + 12) = __ldg(sp + 3);
            PairEval e;
            if (eval_pair<kRef, false>(pixx, pixy, s, e)) alpha = e.alpha;
        }
        // sequential-exact transmittance: T_i = T_{i-1} * (1 - alpha_{i-1}) in list order, same roundings as
        // the forward (a left-to-right product), obtained by passing the running value lane to lane.
        float Tin = T;
        unsigned contrib = __ballot_sync(0xffffffffu, alpha != 0.0f);
        while (contrib) {  // contributing lanes only, in list order
            const int l = __ffs(contrib) - 1;
            contrib &= contrib - 1;
            const float a_l = __shfl_sync(0xffffffffu, alpha, l);
            if (lane == l) Tin = T;
            T = mul(T, sub(1.0f, a_l));
        }
        if (alpha != 0.0f) {
            const float w = mul(alpha, Tin);
            if ((F & 3) == 0) {  // 16-byte vector reductions (red.global.add.v4.f32): 4x fewer atomic operations
                float4* dst = reinterpret_cast<float4*>(dL_dextras + (size_t)g * F);
#pragma unroll
                for (int v4 = 0; v4 < FP / 4; v4++)
                    if (v4 * 4 < F)
                        atomicAdd(dst + v4, make_float4(mul(w, dE[4 * v4]), mul(w, dE[4 * v4 + 1]), mul(w, dE[4 * v4 + 2]),
                                                        mul(w, dE[4 * v4 + 3])));
            } else {
#pragma unroll
                for (int ch = 0; ch < FP; ch++)
                    if (ch < F) {
                        const float val = mul(w, dE[ch]);
                        if (val != 0.0f) atomicAdd(dL_dextras + (size_t)g * F + ch, val);
                    }
            }
        }
    }
}

template <int FP>
static int launch_bwd_one(const IsrBackwardArgs& a, cudaStream_t stream) {
    GeomLayout gl(a.P);
    ImageLayout il(a.W, a.H);
    const char* g = static_cast<const char*>(a.geom);
    const char* im = static_cast<const char*>(a.image);
    const char* b = static_cast<const char*>(a.binning);
    const int num_tiles = ((a.W + TILE - 1) / TILE) * ((a.H + TILE - 1) / TILE);
    const unsigned m = a.grad_mask;
    auto launch = [&](auto kern, int w) -> int {
        const size_t smem = (size_t)w * BwdSmem<FP>::per_warp;
        ISR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<num_tiles * (8 / w), 32 * w, smem, stream>>>(
            reinterpret_cast<const uint2*>(im + il.ranges), reinterpret_cast<const uint32_t*>(b), a.W, a.H, a.F, a.background,
            reinterpret_cast<const float4*>(g + gl.splat), reinterpret_cast<const float4*>(g + gl.cull),
            reinterpret_cast<const float4*>(g + gl.cullq), reinterpret_cast<const float4*>(g + gl.rgb), a.extra_attrs,
            reinterpret_cast<const float*>(im + il.final_T), reinterpret_cast<const uint32_t*>(im + il.n_contrib), a.dL_dcolor,
            a.dL_dothers, a.dL_dextra_pix, (m & ISR_GRAD_GEOMETRY) ? a.dL_dtransMat : nullptr,
            (m & ISR_GRAD_GEOMETRY) ? a.dL_dmeans2D : nullptr, (m & ISR_GRAD_GEOMETRY) ? a.dL_dnormal : nullptr,
            (m & ISR_GRAD_OPACITY) ? a.dL_dopacity : nullptr, (m & ISR_GRAD_COLOR) ? a.dL_dcolors : nullptr,
            (m & ISR_GRAD_EXTRA) ? a.dL_dextra : nullptr, entries_packed(a.P) ? 1 : 0); note_launch();
        ISR_CUDA_TRY(cudaGetLastError());
        return ISR_OK;
    };
    // measured at cfg2: 2-warp CTAs are 8% faster than whole-tile CTAs
    if (a.flags & ISR_FLAG_SPEC_ARITH) return launch(blend_bwd_kernel<FP, 2, false>, 2);
    return launch(blend_bwd_kernel<FP, 2, true>, 2);
}

int launch_blend_bwd(const IsrBackwardArgs& a, cudaStream_t stream) {
    if (a.num_rendered <= 0) return ISR_OK;
    const int F = a.F;
    if (F == 0) return launch_bwd_one<0>(a, stream);
    if (F <= 4) return launch_bwd_one<4>(a, stream);
    if (F <= 8) return launch_bwd_one<8>(a, stream);
    if (F <= 16) return launch_bwd_one<16>(a, stream);
    if (F <= 24) return launch_bwd_one<24>(a, stream);
    if (F <= 32) return launch_bwd_one<32>(a, stream);
    return ISR_ERR_UNSUPPORTED;
}

static SparseViewsDev make_views(int n_views, const IsrSparseView* views, int P, int W, int H) {
    SparseViewsDev v{};
    GeomLayout gl(P);
    ImageLayout il(W, H);
    v.n_views = n_views;
    for (int i = 0; i < n_views; i++) {
        const char* g = static_cast<const char*>(views[i].geom);
        char* im = static_cast<char*>(views[i].image);
        v.ranges[i] = reinterpret_cast<const uint2*>(im + il.ranges);
        v.point_list[i] = reinterpret_cast<const uint32_t*>(views[i].binning);  // offset 0 of the binning workspace
        v.splats[i] = reinterpret_cast<const float4*>(g + gl.splat);
        v.cull4[i] = reinterpret_cast<const float4*>(g + gl.cull);
        v.n_contrib[i] = reinterpret_cast<uint32_t*>(im + il.n_contrib);
        v.final_T[i] = reinterpret_cast<float*>(im + il.final_T);
    }
    return v;
}

template <int FP>
static int launch_sparse_bwd_one(const SparseViewsDev& v, int P, int F, int W, int H, int n, const int* pix_ids,
                                 const int* view_ids, const float* dLdE, float* dL_dextra, unsigned flags, cudaStream_t stream) {
    const int warps_per_block = 8;
    auto kern = (flags & ISR_FLAG_SPEC_ARITH) ? extra_sparse_bwd_kernel<FP, false> : extra_sparse_bwd_kernel<FP, true>;
    kern<<<(n + warps_per_block - 1) / warps_per_block, 256, 0, stream>>>(v, n, pix_ids, view_ids, dLdE, W, H, F, dL_dextra,
                                                                        entries_packed(P) ? 1 : 0); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_sparse_bwd_views(int n_views, const IsrSparseView* views, int P, int F, int W, int H, int n, const int* pix_ids,
                            const int* view_ids, const float* dLdE, float* dL_dextra, unsigned flags, cudaStream_t stream) {
    if (n <= 0 || F <= 0) return ISR_OK;
    const SparseViewsDev v = make_views(n_views, views, P, W, H);
#define ISR_SB(FPV) return launch_sparse_bwd_one<FPV>(v, P, F, W, H, n, pix_ids, view_ids, dLdE, dL_dextra, flags, stream)
    if (F <= 4) ISR_SB(4);
    if (F <= 8) ISR_SB(8);
    if (F <= 16) ISR_SB(16);
    if (F <= 24) ISR_SB(24);
    if (F <= 32) ISR_SB(32);
#undef ISR_SB
    return ISR_ERR_UNSUPPORTED;
}

int launch_extra_sparse_bwd(int P, int F, int W, int H, const void* geom, const void* image, const void* binning, int n,
                            const int* pix_ids, const float* dLdE, float* dL_dextra, unsigned flags, cudaStream_t stream) {
    IsrSparseView one{geom, const_cast<void*>(image), binning};
    return launch_sparse_bwd_views(1, &one, P, F, W, H, n, pix_ids, nullptr, dLdE, dL_dextra, flags, stream);
}

template <int FP>
static int launch_sparse_fwd_one(const SparseViewsDev& v, int P, int F, int W, int H, const float* extras, int n,
                                 const int* pix_ids, const int* view_ids, float* out, unsigned flags, cudaStream_t stream) {
    const int warps_per_block = 8;
    const size_t smem = (size_t)warps_per_block * 32 * FP * sizeof(float);
    auto kern = (flags & ISR_FLAG_SPEC_ARITH) ? extra_sparse_fwd_kernel<FP, false> : extra_sparse_fwd_kernel<FP, true>;
    ISR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(n + warps_per_block - 1) / warps_per_block, 256, smem, stream>>>(v, n, pix_ids, view_ids, W, H, F, extras, out,
                                                                           entries_packed(P) ? 1 : 0); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_sparse_fwd_views(int n_views, const IsrSparseView* views, int P, int F, int W, int H, const float* extras, int n,
                            const int* pix_ids, const int* view_ids, float* out, unsigned flags, cudaStream_t stream) {
    if (n <= 0 || F <= 0) return ISR_OK;
    const SparseViewsDev v = make_views(n_views, views, P, W, H);
#define ISR_SF(FPV) return launch_sparse_fwd_one<FPV>(v, P, F, W, H, extras, n, pix_ids, view_ids, out, flags, stream)
    if (F <= 4) ISR_SF(4);
    if (F <= 8) ISR_SF(8);
    if (F <= 16) ISR_SF(16);
    if (F <= 24) ISR_SF(24);
    if (F <= 32) ISR_SF(32);
#undef ISR_SF
    return ISR_ERR_UNSUPPORTED;
}

}  // namespace isr
