// isr_preprocess.cu -- per-Gaussian kernels: K1 forward projection (DSR/cuda_rasterizer/forward.cu:148-251),
// K8 backward chain rule (backward.cu:469-656, 20-139) and markVisible (rasterizer_impl.cu:54-66).
// One thread per Gaussian; HBM-streaming kernels (algorithmic bytes in DESIGN.md).
#include "isr_common.cuh"

namespace isr {

// DSR/cuda_rasterizer/auxiliary.h:44-61
__device__ constexpr float SH_C0 = 0.28209479177387814f;
__device__ constexpr float SH_C1 = 0.4886025119029199f;
__device__ constexpr float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                       -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                       0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                       -0.5900435899266435f};

// Camera matrices stay in device memory (they are torch CUDA tensors in the reference API); every thread
// reads the same 35 floats through the read-only path (uniform addresses -> one broadcast transaction).
struct Camera {
    const float* __restrict__ view;    // [16]
    const float* __restrict__ proj;    // [16]
    const float* __restrict__ campos;  // [3]
};

// quat (w,x,y,z) -> rotation columns (auxiliary.h:214-236); rsqrtf per kRef (isr_common.cuh)
template <bool kRef>
__device__ __forceinline__ void quat_to_rot(const float4 q, float R0[3], float R1[3], float R2[3]) {
    const float sum = fma_(q.z, q.z, fma_(q.y, q.y, fma_(q.x, q.x, mul(q.w, q.w))));
    const float s = rsqrt_sel<kRef>(sum);
    const float w = mul(q.x, s), x = mul(q.y, s), y = mul(q.z, s), z = mul(q.w, s);
    const float yy = mul(y, y), zz = mul(z, z);
    const float yy_zz = add(yy, zz), xx_zz = fma_(x, x, zz), xx_yy = fma_(x, x, yy);
    const float xy_p = fma_(x, y, mul(w, z)), xy_m = fma_(x, y, -mul(w, z));
    const float xz_p = fma_(x, z, mul(w, y)), xz_m = fma_(x, z, -mul(w, y));
    const float yz_p = fma_(y, z, mul(w, x)), yz_m = fma_(y, z, -mul(w, x));
    R0[0] = sub(1.0f, add(yy_zz, yy_zz)); R0[1] = add(xy_p, xy_p);             R0[2] = add(xz_m, xz_m);
    R1[0] = add(xy_m, xy_m);             R1[1] = sub(1.0f, add(xx_zz, xx_zz)); R1[2] = add(yz_p, yz_p);
    R2[0] = add(xz_p, xz_p);             R2[1] = add(yz_m, yz_m);             R2[2] = sub(1.0f, add(xx_yy, xx_yy));
}

// SH staging, variant kBulk: ONE bulk-async copy (cp.async.bulk, the TMA unit's 1-D mode) of the block's 256 contiguous
// rows (48 KB), completion counted in bytes on an mbarrier, instead of twelve 16-byte cp.async per thread.
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(phase)
        : "memory");
}

template <bool kRef, bool kBulk>
__global__ void __launch_bounds__(256)
preprocess_fwd_kernel(int P, int D, int M, const float* __restrict__ means3D, const float2* __restrict__ scales,
                      float scale_modifier, const float4* __restrict__ rotations,
                      const float* __restrict__ opacities, const float* __restrict__ shs,
                      const float* __restrict__ transMat_precomp, const float* __restrict__ colors_precomp,
                      const Camera cam, int W, int H, int gx, int gy, int* __restrict__ radii,
                      Splat* __restrict__ splats, float4* __restrict__ cull4, float4* __restrict__ cullq,
                      float4* __restrict__ rgb4,
                      float* __restrict__ depths,
                      uint32_t* __restrict__ depth_keys, uint32_t* __restrict__ tiles_touched,
                      uint8_t* __restrict__ clamped, uint4* __restrict__ tile_foot,
                      uint32_t* __restrict__ tile_count) {
    // SH coefficients of the block's 256 Gaussians are one contiguous 48 KB range: stage them with fully coalesced
    // 16-byte cp.async copies into per-Gaussian slots padded to 13 x 16 B (conflict-free LDS), overlapped with the
    // projection math below.  (Per-thread strided loads of the 192-byte rows ran K1 at 48% of the HBM roofline.)
    extern __shared__ __align__(16) float4 s_sh[];
    __shared__ __align__(8) uint64_t s_bar;
    const bool stage_sh = (shs != nullptr) && (M == 16) && (colors_precomp == nullptr);
    if (stage_sh) {
        const int block_base = blockIdx.x * blockDim.x;
        const int n_rows = min((int)blockDim.x, P - block_base);
        if constexpr (kBulk) {
            if (threadIdx.x == 0) mbar_init(&s_bar, 1);
            __syncthreads();
            // the block's rows are one contiguous range: ONE copy of up to 48 KB issued by one thread (UBLKCP takes uniform
            // operands: a copy per row would be serialised lane by lane), landing unpadded (row stride 192 B)
            if (threadIdx.x == 0) {
                mbar_arrive_expect_tx(&s_bar, (unsigned)n_rows * 192u);
                bulk_copy_g2s(s_sh, shs + (size_t)block_base * 48, (unsigned)n_rows * 192u, &s_bar);
            }
        } else {
            const int n_chunks = n_rows * 12;
            const float4* src = reinterpret_cast<const float4*>(shs + (size_t)block_base * 48);
            for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) {
                const int row = c / 12, col = c - row * 12;
                const unsigned d = (unsigned)__cvta_generic_to_shared(s_sh + row * 13 + col);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + c));
            }
            asm volatile("cp.async.commit_group;\n" ::);
        }
    }
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int radius_i = 0;
    uint32_t tiles = 0, dkey = 0xFFFFFFFFu, tcount = 0;
    unsigned long long tmask = 0ull;
    uint2 trect = make_uint2(0u, 1u);
    uint8_t clamp_mask = 0;
    bool visible = false;
    float p0 = 0, p1 = 0, p2 = 0, pvz = 0, cx = 0, cy = 0;
    float T[9], nrm[3];
    int ri = 0, ntiles = 0;
    do {
        if (idx >= P) break;
        p0 = means3D[3 * (size_t)idx]; p1 = means3D[3 * (size_t)idx + 1]; p2 = means3D[3 * (size_t)idx + 2];
        float view[16], proj[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { view[i] = __ldg(cam.view + i); proj[i] = __ldg(cam.proj + i); }
        const float pvx = add(dot3c(view[0], p0, view[4], p1, view[8], p2), view[12]);
        const float pvy = add(dot3c(view[1], p0, view[5], p1, view[9], p2), view[13]);
        pvz = add(dot3c(view[2], p0, view[6], p1, view[10], p2), view[14]);
        if (pvz <= 0.2f) break;
        if (transMat_precomp == nullptr) {
            const float2 sc = scales[idx];
            const float sx = mul(scale_modifier, sc.x), sy = mul(scale_modifier, sc.y);
            float R0[3], R1[3], R2[3];
            quat_to_rot<kRef>(rotations[idx], R0, R1, R2);
            const float L0[3] = {mul(R0[0], sx), mul(R0[1], sx), mul(R0[2], sx)};
            const float L1[3] = {mul(R1[0], sy), mul(R1[1], sy), mul(R1[2], sy)};
            float X[4][3];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                X[c][0] = dot3c(L0[0], proj[c], L0[1], proj[4 + c], L0[2], proj[8 + c]);
                X[c][1] = dot3c(L1[0], proj[c], L1[1], proj[4 + c], L1[2], proj[8 + c]);
                X[c][2] = add(dot3c(p0, proj[c], p1, proj[4 + c], p2, proj[8 + c]), proj[12 + c]);
            }
            const float hw = mul((float)W, 0.5f), hw1 = mul((float)(W - 1), 0.5f);
            const float hh = mul((float)H, 0.5f), hh1 = mul((float)(H - 1), 0.5f);
#pragma unroll
            for (int r = 0; r < 3; r++) {
                T[0 + r] = fma_(X[3][r], hw1, mul(X[0][r], hw));
                T[3 + r] = fma_(X[3][r], hh1, mul(X[1][r], hh));
                T[6 + r] = X[3][r];
            }
            nrm[0] = dot3c(view[0], R2[0], view[4], R2[1], view[8], R2[2]);
            nrm[1] = dot3c(view[1], R2[0], view[5], R2[1], view[9], R2[2]);
            nrm[2] = dot3c(view[2], R2[0], view[6], R2[1], view[10], R2[2]);
        } else {
#pragma unroll
            for (int i = 0; i < 9; i++) T[i] = transMat_precomp[9 * (size_t)idx + i];
            nrm[0] = 0.0f; nrm[1] = 0.0f; nrm[2] = 1.0f;
        }
        const float cosv = -fma_(pvz, nrm[2], fma_(pvx, nrm[0], mul(pvy, nrm[1])));
        if (cosv == 0.0f) break;
        const float mult = cosv > 0.0f ? 1.0f : -1.0f;
        nrm[0] = mul(nrm[0], mult); nrm[1] = mul(nrm[1], mult); nrm[2] = mul(nrm[2], mult);
        // compute_aabb (forward.cu:119-145), cutoff 3
        const float* Tu = T; const float* Tv = T + 3; const float* Tw = T + 6;
        const float d = fma_(-Tw[2], Tw[2], fma_(mul(Tw[0], Tw[0]), 9.0f, mul(mul(Tw[1], Tw[1]), 9.0f)));
        if (d == 0.0f) break;
        const float inv_d = rcp(d);
        const float f9 = mul(inv_d, 9.0f);
        // [sass] every dot(f, a*b): fma(a.z*b.z, -+inv_d, fma(f9, a.x*b.x, f9*(a.y*b.y))) -- the FMA carries the x term
        cx = fma_(mul(Tu[2], Tw[2]), -inv_d, fma_(f9, mul(Tu[0], Tw[0]), mul(f9, mul(Tu[1], Tw[1]))));
        cy = fma_(mul(Tv[2], Tw[2]), -inv_d, fma_(f9, mul(Tv[0], Tw[0]), mul(f9, mul(Tv[1], Tw[1]))));
        const float ngx = fma_(mul(Tu[2], Tu[2]), inv_d, -fma_(f9, mul(Tu[0], Tu[0]), mul(f9, mul(Tu[1], Tu[1]))));
        const float ngy = fma_(mul(Tv[2], Tv[2]), inv_d, -fma_(f9, mul(Tv[0], Tv[0]), mul(f9, mul(Tv[1], Tv[1]))));
        const float ex = sqrt_(fmaxf(1e-4f, fma_(cx, cx, ngx)));
        const float ey = sqrt_(fmaxf(1e-4f, fma_(cy, cy, ngy)));
        const float radius = ceilf(fmaxf(fmaxf(ex, ey), mul(3.0f, kFilterSize)));
        ri = __float2int_rz(radius);
        int mnx, mny, mxx, mxy;
        get_rect(cx, cy, ri, gx, gy, mnx, mny, mxx, mxy);
        ntiles = (mxx - mnx) * (mxy - mny);
        if (ntiles == 0) break;
        visible = true;
    } while (false);
    if (stage_sh) {
        if constexpr (kBulk) {
            mbar_wait(&s_bar, 0);  // every row of the block has landed (each thread reads only its own)
        } else {
            asm volatile("cp.async.wait_all;\n" ::: "memory");
            __syncthreads();
        }
    }
    if (visible) {
        const float* Tu = T; const float* Tv = T + 3; const float* Tw = T + 6;
        float rgb[3];
        if (colors_precomp == nullptr) {
            // computeColorFromSH (forward.cu:20-71)
            const float* sh = stage_sh ? reinterpret_cast<const float*>(s_sh + threadIdx.x * (kBulk ? 12 : 13))
                                       : shs + (size_t)idx * M * 3;
            const float dx = sub(p0, __ldg(cam.campos)), dy = sub(p1, __ldg(cam.campos + 1)), dz = sub(p2, __ldg(cam.campos + 2));
            const float len = sqrt_(fma_(dz, dz, fma_(dx, dx, mul(dy, dy))));
            const float x = __fdiv_rn(dx, len), y = __fdiv_rn(dy, len), z = __fdiv_rn(dz, len);
            float res[3];
#pragma unroll
            for (int c = 0; c < 3; c++) res[c] = mul(SH_C0, sh[c]);
            if (D > 0) {
                const float a1 = mul(SH_C1, y), a2 = mul(SH_C1, z), a3 = mul(SH_C1, x);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float r = res[c];
                    r = fma_(-a1, sh[3 + c], r);
                    r = fma_(a2, sh[6 + c], r);
                    r = fma_(-a3, sh[9 + c], r);
                    res[c] = r;
                }
                if (D > 1) {
                    const float xx = mul(x, x), yy = mul(y, y), zz = mul(z, z);
                    const float xy = mul(x, y), yz = mul(y, z), xz = mul(x, z);
                    const float b4 = mul(SH_C2[0], xy), b5 = mul(SH_C2[1], yz);
                    const float b6 = mul(SH_C2[2], sub(sub(mul(2.0f, zz), xx), yy));
                    const float b7 = mul(SH_C2[3], xz), b8 = mul(SH_C2[4], sub(xx, yy));
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        float r = res[c];
                        r = fma_(b4, sh[12 + c], r);
                        r = fma_(b5, sh[15 + c], r);
                        r = fma_(b6, sh[18 + c], r);
                        r = fma_(b7, sh[21 + c], r);
                        r = fma_(b8, sh[24 + c], r);
                        res[c] = r;
                    }
                    if (D > 2) {
                        // [sass] 3xx-yy = fma(xx,3,-yy); 4zz-xx-yy = fma(zz,4,-xx) - yy (shared by c11, c13);
                        // 2zz-3xx-3yy = fma(yy,-3, fma(xx,-3, zz+zz)); xx-3yy = fma(yy,-3,xx)
                        const float t4 = sub(fma_(zz, 4.0f, -xx), yy);
                        const float c9 = mul(mul(SH_C3[0], y), fma_(xx, 3.0f, -yy));
                        const float c10 = mul(mul(SH_C3[1], xy), z);
                        const float c11 = mul(mul(SH_C3[2], y), t4);
                        const float c12 = mul(mul(SH_C3[3], z), fma_(yy, -3.0f, fma_(xx, -3.0f, add(zz, zz))));
                        const float c13 = mul(mul(SH_C3[4], x), t4);
                        const float c14 = mul(mul(SH_C3[5], z), sub(xx, yy));
                        const float c15 = mul(mul(SH_C3[6], x), fma_(yy, -3.0f, xx));
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            float r = res[c];
                            r = fma_(c9, sh[27 + c], r);
                            r = fma_(c10, sh[30 + c], r);
                            r = fma_(c11, sh[33 + c], r);
                            r = fma_(c12, sh[36 + c], r);
                            r = fma_(c13, sh[39 + c], r);
                            r = fma_(c14, sh[42 + c], r);
                            r = fma_(c15, sh[45 + c], r);
                            res[c] = r;
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float r = add(res[c], 0.5f);
                if (r < 0.0f) clamp_mask |= (uint8_t)(1u << c);
                rgb[c] = fmaxf(r, 0.0f);
            }
        } else {
            rgb[0] = colors_precomp[3 * (size_t)idx];
            rgb[1] = colors_precomp[3 * (size_t)idx + 1];
            rgb[2] = colors_precomp[3 * (size_t)idx + 2];
        }

        const float opa = opacities[idx];
        // conservative alpha cut: alpha = opa*exp(power) < 1/255  <=  power < -ln(255*opa) - margin
        float power_cut = __int_as_float(0x7f800000);  // +inf: never contributes (opa <= 0)
        // (never below -80: there the specified exp is 0, so alpha = 0 and the pair is skipped anyway)
        if (opa > 0.0f) power_cut = fmaxf(-__logf(255.0f * opa) - 1e-3f, -80.0f);
        Splat s;
        s.Tu[0] = T[0]; s.Tu[1] = T[1]; s.Tu[2] = T[2];
        s.Tv[0] = T[3]; s.Tv[1] = T[4]; s.Tv[2] = T[5];
        s.Tw[0] = T[6]; s.Tw[1] = T[7]; s.Tw[2] = T[8];
        s.mx = cx; s.my = cy;
        s.nx = nrm[0]; s.ny = nrm[1]; s.nz = nrm[2];
        s.opacity = opa;
        s.power_cut = power_cut;
        float4* dst = reinterpret_cast<float4*>(splats + idx);
        const float4* src = reinterpret_cast<const float4*>(&s);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        // Conservative screen-space rectangle outside of which this Gaussian cannot pass the alpha test:
        // a pair survives only if min(rho3d, rho2d) <= rho_max = -2*power_cut.  rho3d <= rho_max is the
        // projection of the disk u^2+v^2 <= rho_max (an ellipse when dd < 0, same algebra as compute_aabb);
        // rho2d <= rho_max is a disk of radius sqrt(rho_max/2) around the AABB centre.  Margins absorb fp32
        // rounding; non-elliptic projections disable the test.  Used only to SKIP work, never to change results.
        float4 cr = make_float4(1e30f, 1e30f, -1e30f, -1e30f);  // empty
        const float rho_max = -2.0f * power_cut;
        if (rho_max > 0.0f) {
            const float INF = __int_as_float(0x7f800000);
            const float tw2 = Tw[2] * Tw[2];
            const float dd = rho_max * (Tw[0] * Tw[0] + Tw[1] * Tw[1]) - tw2;
            if (!(dd < -1e-3f * tw2)) {
                cr = make_float4(-INF, -INF, INF, INF);
            } else {
                const float fa = rho_max / dd, fz = -1.0f / dd;
                const float ccx = fa * (Tu[0] * Tw[0] + Tu[1] * Tw[1]) + fz * Tu[2] * Tw[2];
                const float ccy = fa * (Tv[0] * Tw[0] + Tv[1] * Tw[1]) + fz * Tv[2] * Tw[2];
                const float hx2 = ccx * ccx - (fa * (Tu[0] * Tu[0] + Tu[1] * Tu[1]) + fz * Tu[2] * Tu[2]);
                const float hy2 = ccy * ccy - (fa * (Tv[0] * Tv[0] + Tv[1] * Tv[1]) + fz * Tv[2] * Tv[2]);
                const float hx = sqrtf(fmaxf(hx2, 0.0f)), hy = sqrtf(fmaxf(hy2, 0.0f));
                const float mgx = 0.5f + 1e-3f * (hx + fabsf(ccx)), mgy = 0.5f + 1e-3f * (hy + fabsf(ccy));
                const float r2 = sqrtf(0.5f * rho_max) + 0.5f;
                cr.x = fminf(ccx - hx - mgx, cx - r2);
                cr.y = fminf(ccy - hy - mgy, cy - r2);
                cr.z = fmaxf(ccx + hx + mgx, cx + r2);
                cr.w = fmaxf(ccy + hy + mgy, cy + r2);
                if (!(hx2 == hx2) || !(hy2 == hy2)) cr = make_float4(-INF, -INF, INF, INF);
            }
        }
        cull4[idx] = cr;
        // Normalised conic of the same region for the second-stage block test (isr::block_outside), in fp64:
        // p(x,y) = A x + B y + C with A = Tv x Tw, B = Tw x Tu, C = Tu x Tv (k x l expanded), re-centred at (cx,cy).
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = make_float4(0.f, -1.f, cx, cy);  // Q == -1: never "outside"
        float r2 = __int_as_float(0x7f800000);
        if (rho_max > 0.0f && cr.x > -1e29f) {
            const double rho = (double)rho_max * 1.0002;
            const double tu0 = Tu[0], tu1 = Tu[1], tu2 = Tu[2], tv0 = Tv[0], tv1 = Tv[1], tv2 = Tv[2];
            const double tw0 = Tw[0], tw1 = Tw[1], tw2 = Tw[2];
            const double Ax = tv1 * tw2 - tv2 * tw1, Ay = tv2 * tw0 - tv0 * tw2, Az = tv0 * tw1 - tv1 * tw0;
            const double Bx = tw1 * tu2 - tw2 * tu1, By = tw2 * tu0 - tw0 * tu2, Bz = tw0 * tu1 - tw1 * tu0;
            const double Cx = tu1 * tv2 - tu2 * tv1, Cy = tu2 * tv0 - tu0 * tv2, Cz = tu0 * tv1 - tu1 * tv0;
            const double mx = cx, my = cy;
            const double cpx = Ax * mx + Bx * my + Cx, cpy = Ay * mx + By * my + Cy, cpz = Az * mx + Bz * my + Cz;
            const double qxx = Ax * Ax + Ay * Ay - rho * Az * Az, qyy = Bx * Bx + By * By - rho * Bz * Bz;
            const double qxy = 2.0 * (Ax * Bx + Ay * By - rho * Az * Bz);
            const double qx = 2.0 * (Ax * cpx + Ay * cpy - rho * Az * cpz), qy = 2.0 * (Bx * cpx + By * cpy - rho * Bz * cpz);
            const double qc = cpx * cpx + cpy * cpy - rho * cpz * cpz;
            const double nrm2 = rho * cpz * cpz;
            // usable only for a proper ellipse that contains its own centre
            if (nrm2 > 0.0 && qc < 0.0 && qxx > 0.0 && 4.0 * qxx * qyy - qxy * qxy > 0.0) {
                const double sc = 1.0 / nrm2;
                q0 = make_float4((float)(qxx * sc), (float)(qxy * sc), (float)(qyy * sc), (float)(qx * sc));
                q1 = make_float4((float)(qy * sc), (float)(qc * sc), cx, cy);
            }
            const float rd = sqrtf(0.5f * rho_max) + 0.5f;
            r2 = rd * rd;
        } else if (!(rho_max > 0.0f)) {
            q1.y = 1e30f;  // never contributes: Q huge, disk empty
            r2 = -1.0f;
        }
        cullq[3 * (size_t)idx + 0] = q0;
        cullq[3 * (size_t)idx + 1] = q1;
        cullq[3 * (size_t)idx + 2] = make_float4(r2, 0.f, 0.f, 0.f);
        // Tile footprint for the binning (isr_binning.cu): bit t of the mask = tile t (row-major) of the reference's
        // getRect rectangle may receive something from this Gaussian; the others are never emitted (skipping them
        // cannot change a result).  Rectangles of more than 64 tiles are emitted whole.
        int mnx, mny, mxx, mxy;
        get_rect(cx, cy, ri, gx, gy, mnx, mny, mxx, mxy);
        const int w = mxx - mnx;
        trect = make_uint2((uint32_t)mnx | ((uint32_t)mny << 16), (uint32_t)w);  // getRect origin, width (binning)
        if (ntiles <= 64) {
            // only the tiles overlapping the cull rectangle can pass (a superset is harmless, hence the slack)
            const int tx_lo = max(mnx, __float2int_ru((cr.x - 15.0f) * 0.0625f - 1e-3f));
            const int tx_hi = min(mxx - 1, __float2int_rd(cr.z * 0.0625f + 1e-3f));
            const int ty_lo = max(mny, __float2int_ru((cr.y - 15.0f) * 0.0625f - 1e-3f));
            const int ty_hi = min(mxy - 1, __float2int_rd(cr.w * 0.0625f + 1e-3f));
            for (int ty = ty_lo; ty <= ty_hi; ty++)
                for (int tx = tx_lo; tx <= tx_hi; tx++)
                    if (rect_may_touch(tx * TILE, ty * TILE, TILE, TILE, W, H, cr, q0, q1, r2))
                        tmask |= 1ull << ((ty - mny) * w + (tx - mnx));
            tcount = (uint32_t)__popcll(tmask);
        } else {
            tmask = ~0ull;
            tcount = (uint32_t)ntiles;
        }
        rgb4[idx] = make_float4(rgb[0], rgb[1], rgb[2], 0.0f);
        depths[idx] = pvz;
        dkey = __float_as_uint(pvz);
        radius_i = ri;
        tiles = (uint32_t)ntiles;
    }
    if (idx >= P) return;
    radii[idx] = radius_i;
    tiles_touched[idx] = tiles;
    depth_keys[idx] = dkey;
    clamped[idx] = clamp_mask;
    tile_count[idx] = tcount;
    tile_foot[2 * (size_t)idx] = make_uint4(tcount | (tcount > 64u ? 0x80000000u : 0u), trect.x, trect.y, 0u);
    tile_foot[2 * (size_t)idx + 1] = make_uint4((uint32_t)tmask, (uint32_t)(tmask >> 32), 0u, 0u);
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const Camera cam,
                                    uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float p0 = means3D[3 * (size_t)idx], p1 = means3D[3 * (size_t)idx + 1], p2 = means3D[3 * (size_t)idx + 2];
    const float pvz = add(dot3c(__ldg(cam.view + 2), p0, __ldg(cam.view + 6), p1, __ldg(cam.view + 10), p2), __ldg(cam.view + 14));
    present[idx] = pvz > 0.2f;
}

// ---------------------------------------------------------------------------------------------------------
// K8: backward of the per-Gaussian projection (backward.cu:469-656) + SH backward (backward.cu:20-139).
// Float-only results; written as plain expressions (tolerance-compared with the oracle / reference).
// ---------------------------------------------------------------------------------------------------------
// SH backward (reference: backward.cu:20-139) written as "basis + basis gradient":
//   rgb = sum_i Y_i(d) * sh_i + 0.5,  d = (pos - campos)/|pos - campos|
//   dL/dsh_i = Y_i * dL/drgb (clamped channels masked),   dL/dd = sum_i gradY_i * (sh_i . dL/drgb)
// followed by the Jacobian of the normalisation (auxiliary.h:129-139).
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* Y, float* Yx, float* Yy, float* Yz) {
    for (int i = 0; i < 16; i++) { Y[i] = 0.0f; Yx[i] = 0.0f; Yy[i] = 0.0f; Yz[i] = 0.0f; }
    Y[0] = SH_C0;
    if (deg < 1) return;
    Y[1] = -SH_C1 * y; Yy[1] = -SH_C1;
    Y[2] = SH_C1 * z;  Yz[2] = SH_C1;
    Y[3] = -SH_C1 * x; Yx[3] = -SH_C1;
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    Y[4] = SH_C2[0] * xy;                   Yx[4] = SH_C2[0] * y;          Yy[4] = SH_C2[0] * x;
    Y[5] = SH_C2[1] * yz;                   Yy[5] = SH_C2[1] * z;          Yz[5] = SH_C2[1] * y;
    Y[6] = SH_C2[2] * (2.f * zz - xx - yy); Yx[6] = -2.f * SH_C2[2] * x;   Yy[6] = -2.f * SH_C2[2] * y;  Yz[6] = 4.f * SH_C2[2] * z;
    Y[7] = SH_C2[3] * xz;                   Yx[7] = SH_C2[3] * z;          Yz[7] = SH_C2[3] * x;
    Y[8] = SH_C2[4] * (xx - yy);            Yx[8] = 2.f * SH_C2[4] * x;    Yy[8] = -2.f * SH_C2[4] * y;
    if (deg < 3) return;
    Y[9] = SH_C3[0] * y * (3.f * xx - yy);
    Yx[9] = SH_C3[0] * 6.f * xy;            Yy[9] = SH_C3[0] * 3.f * (xx - yy);
    Y[10] = SH_C3[1] * xy * z;
    Yx[10] = SH_C3[1] * yz;                 Yy[10] = SH_C3[1] * xz;        Yz[10] = SH_C3[1] * xy;
    Y[11] = SH_C3[2] * y * (4.f * zz - xx - yy);
    Yx[11] = SH_C3[2] * -2.f * xy;          Yy[11] = SH_C3[2] * (4.f * zz - xx - 3.f * yy);  Yz[11] = SH_C3[2] * 8.f * yz;
    Y[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
    Yx[12] = SH_C3[3] * -6.f * xz;          Yy[12] = SH_C3[3] * -6.f * yz; Yz[12] = SH_C3[3] * 3.f * (2.f * zz - xx - yy);
    Y[13] = SH_C3[4] * x * (4.f * zz - xx - yy);
    Yx[13] = SH_C3[4] * (4.f * zz - 3.f * xx - yy);  Yy[13] = SH_C3[4] * -2.f * xy;  Yz[13] = SH_C3[4] * 8.f * xz;
    Y[14] = SH_C3[5] * z * (xx - yy);
    Yx[14] = SH_C3[5] * 2.f * xz;           Yy[14] = SH_C3[5] * -2.f * yz; Yz[14] = SH_C3[5] * (xx - yy);
    Y[15] = SH_C3[6] * x * (xx - 3.f * yy);
    Yx[15] = SH_C3[6] * 3.f * (xx - yy);    Yy[15] = SH_C3[6] * -6.f * xy;
}

__device__ void sh_backward(int idx, int deg, int M, const float* __restrict__ means, const float* campos,
                            const float* __restrict__ shs, uint8_t clamp_mask, const float* __restrict__ dL_dcolor,
                            float* __restrict__ dL_dmeans, float* __restrict__ dL_dshs) {
    const float ox = means[3 * (size_t)idx] - __ldg(campos), oy = means[3 * (size_t)idx + 1] - __ldg(campos + 1),
                oz = means[3 * (size_t)idx + 2] - __ldg(campos + 2);
    const float len2 = ox * ox + oy * oy + oz * oz;
    const float inv_len = 1.0f / sqrtf(len2);
    float Y[16], Yx[16], Yy[16], Yz[16];
    sh_basis(deg, ox * inv_len, oy * inv_len, oz * inv_len, Y, Yx, Yy, Yz);
    float g[3];  // dL/drgb with the clamp mask applied (PyTorch clamp rule: zero gradient where clamped)
    for (int c = 0; c < 3; c++) g[c] = ((clamp_mask >> c) & 1) ? 0.0f : dL_dcolor[3 * (size_t)idx + c];
    const float* sh = shs + (size_t)idx * M * 3;
    float* dsh = dL_dshs + (size_t)idx * M * 3;
    const int n = (deg + 1) * (deg + 1);
    float ddx = 0.0f, ddy = 0.0f, ddz = 0.0f;  // dL/d(direction)
    for (int i = 0; i < n; i++) {
        const float proj = sh[3 * i] * g[0] + sh[3 * i + 1] * g[1] + sh[3 * i + 2] * g[2];
        ddx += Yx[i] * proj; ddy += Yy[i] * proj; ddz += Yz[i] * proj;
        dsh[3 * i] = Y[i] * g[0]; dsh[3 * i + 1] = Y[i] * g[1]; dsh[3 * i + 2] = Y[i] * g[2];
    }
    // d(v/|v|)/dv applied to (ddx, ddy, ddz):  (|v|^2 I - v v^T) dd / |v|^3
    const float inv3 = inv_len * inv_len * inv_len;
    const float vd = ox * ddx + oy * ddy + oz * ddz;
    dL_dmeans[3 * (size_t)idx + 0] += (len2 * ddx - ox * vd) * inv3;
    dL_dmeans[3 * (size_t)idx + 1] += (len2 * ddy - oy * vd) * inv3;
    dL_dmeans[3 * (size_t)idx + 2] += (len2 * ddz - oz * vd) * inv3;
}

template <bool kRef>
__global__ void __launch_bounds__(256)
preprocess_bwd_kernel(int P, int D, int M, const float* __restrict__ means3D, const int* __restrict__ radii,
                      const float* __restrict__ shs, const uint8_t* __restrict__ clamped,
                      const float2* __restrict__ scales, const float4* __restrict__ rotations,
                      const float* __restrict__ transMat_precomp, const Splat* __restrict__ splats,
                      const Camera cam, int W, int H, float* __restrict__ dL_dmean2D,
                      const float* __restrict__ dL_dnormal3D, float* __restrict__ dL_dtransMat,
                      const float* __restrict__ dL_dcolors, float* __restrict__ dL_dsh,
                      float* __restrict__ dL_dmean3D, float* __restrict__ dL_dscales, float* __restrict__ dL_drots) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P || !(radii[idx] > 0)) return;
    const bool precomp = (scales == nullptr);
    float view[16], proj[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { view[i] = __ldg(cam.view + i); proj[i] = __ldg(cam.proj + i); }
    float T[3][3], Pm[3][4], R[3][3], normal[3] = {0, 0, 0};
    float sx = 0, sy = 0;
    float4 q = make_float4(1, 0, 0, 0);
    const float p[3] = {means3D[3 * (size_t)idx], means3D[3 * (size_t)idx + 1], means3D[3 * (size_t)idx + 2]};
    if (precomp) {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[i][j] = transMat_precomp[9 * (size_t)idx + 3 * i + j];
    } else {
        q = rotations[idx];
        quat_to_rot<kRef>(q, R[0], R[1], R[2]);
        const float2 sc = scales[idx];
        sx = sc.x; sy = sc.y;  // Q6: scale_modifier ignored (backward.cu:507)
        const float L0[3] = {R[0][0] * sx, R[0][1] * sx, R[0][2] * sx};
        const float L1[3] = {R[1][0] * sy, R[1][1] * sy, R[1][2] * sy};
        const float Mm[3][4] = {{L0[0], L0[1], L0[2], 0.0f}, {L1[0], L1[1], L1[2], 0.0f}, {p[0], p[1], p[2], 1.0f}};
        const float n2p[3][4] = {{(float)W * 0.5f, 0, 0, (float)(W - 1) * 0.5f},
                                 {0, (float)H * 0.5f, 0, (float)(H - 1) * 0.5f},
                                 {0, 0, 0, 1.0f}};
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 4; r++) {
                float s = 0;
                for (int k = 0; k < 4; k++) s += proj[4 * r + k] * n2p[c][k];
                Pm[c][r] = s;
            }
        for (int c = 0; c < 3; c++)
            for (int i = 0; i < 3; i++) {
                float s = 0;
                for (int k = 0; k < 4; k++) s += Mm[i][k] * Pm[c][k];
                T[c][i] = s;
            }
        normal[0] = view[0] * R[2][0] + view[4] * R[2][1] + view[8] * R[2][2];
        normal[1] = view[1] * R[2][0] + view[5] * R[2][1] + view[9] * R[2][2];
        normal[2] = view[2] * R[2][0] + view[6] * R[2][1] + view[10] * R[2][2];
    }
    float dT[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dT[i][j] = dL_dtransMat[9 * (size_t)idx + 3 * i + j];
    const float dT_u2 = dT[0][2], dT_v2 = dT[1][2];
    const float dm2x = dL_dmean2D[3 * (size_t)idx], dm2y = dL_dmean2D[3 * (size_t)idx + 1];
    bool wrote_back = false;
    if (dm2x != 0 || dm2y != 0) {
        const float tv[3] = {9.0f, 9.0f, -1.0f};
        const float d = tv[0] * T[2][0] * T[2][0] + tv[1] * T[2][1] * T[2][1] + tv[2] * T[2][2] * T[2][2];
        float fv[3], dT3[3], df[3];
        for (int j = 0; j < 3; j++) fv[j] = tv[j] * (1.0f / d);
        for (int j = 0; j < 3; j++) {
            dT[0][j] += dm2x * fv[j] * T[2][j];
            dT[1][j] += dm2y * fv[j] * T[2][j];
            dT3[j] = dm2x * fv[j] * T[0][j] + dm2y * fv[j] * T[1][j];
            df[j] = dm2x * T[0][j] * T[2][j] + dm2y * T[1][j] * T[2][j];
        }
        const float dL_dd = (df[0] * fv[0] + df[1] * fv[1] + df[2] * fv[2]) * (-1.0f / d);
        for (int j = 0; j < 3; j++) dT[2][j] += dT3[j] + dL_dd * (tv[j] * T[2][j] * 2.0f);
        if (precomp) {
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dL_dtransMat[9 * (size_t)idx + 3 * i + j] = dT[i][j];
            wrote_back = true;
        }
    }
    if (!precomp) {
        float dM[3][4];
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 4; r++) dM[c][r] = Pm[0][r] * dT[0][c] + Pm[1][r] * dT[1][c] + Pm[2][r] * dT[2][c];
        const float* dn = dL_dnormal3D + 3 * (size_t)idx;
        float dtn[3] = {view[0] * dn[0] + view[1] * dn[1] + view[2] * dn[2],
                        view[4] * dn[0] + view[5] * dn[1] + view[6] * dn[2],
                        view[8] * dn[0] + view[9] * dn[1] + view[10] * dn[2]};
        const float pvx = view[0] * p[0] + view[4] * p[1] + view[8] * p[2] + view[12];
        const float pvy = view[1] * p[0] + view[5] * p[1] + view[9] * p[2] + view[13];
        const float pvz = view[2] * p[0] + view[6] * p[1] + view[10] * p[2] + view[14];
        const float cosv = -(pvx * normal[0] + pvy * normal[1] + pvz * normal[2]);
        const float mult = cosv > 0 ? 1.0f : -1.0f;
        for (int j = 0; j < 3; j++) dtn[j] *= mult;
        const float dRS[3][3] = {{dM[0][0], dM[0][1], dM[0][2]}, {dM[1][0], dM[1][1], dM[1][2]}, {dtn[0], dtn[1], dtn[2]}};
        float vR[3][3];
        for (int j = 0; j < 3; j++) { vR[0][j] = dRS[0][j] * sx; vR[1][j] = dRS[1][j] * sy; vR[2][j] = dRS[2][j]; }
        const float s = 1.0f / sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
        const float w = q.x * s, x = q.y * s, y = q.z * s, z = q.w * s;
        float4 dq;
        dq.x = 2.f * (x * (vR[1][2] - vR[2][1]) + y * (vR[2][0] - vR[0][2]) + z * (vR[0][1] - vR[1][0]));
        dq.y = 2.f * (-2.f * x * (vR[1][1] + vR[2][2]) + y * (vR[0][1] + vR[1][0]) + z * (vR[0][2] + vR[2][0]) + w * (vR[1][2] - vR[2][1]));
        dq.z = 2.f * (x * (vR[0][1] + vR[1][0]) - 2.f * y * (vR[0][0] + vR[2][2]) + z * (vR[1][2] + vR[2][1]) + w * (vR[2][0] - vR[0][2]));
        dq.w = 2.f * (x * (vR[0][2] + vR[2][0]) + y * (vR[1][2] + vR[2][1]) - 2.f * z * (vR[0][0] + vR[1][1]) + w * (vR[0][1] - vR[1][0]));
        reinterpret_cast<float4*>(dL_drots)[idx] = dq;
        dL_dscales[2 * (size_t)idx + 0] = dRS[0][0] * R[0][0] + dRS[0][1] * R[0][1] + dRS[0][2] * R[0][2];
        dL_dscales[2 * (size_t)idx + 1] = dRS[1][0] * R[1][0] + dRS[1][1] * R[1][1] + dRS[1][2] * R[1][2];
        dL_dmean3D[3 * (size_t)idx + 0] = dM[2][0];
        dL_dmean3D[3 * (size_t)idx + 1] = dM[2][1];
        dL_dmean3D[3 * (size_t)idx + 2] = dM[2][2];
    }
    if (shs != nullptr && dL_dsh != nullptr)
        sh_backward(idx, D, M, means3D, cam.campos, shs, clamped[idx], dL_dcolors, dL_dmean3D, dL_dsh);
    // densification proxy (backward.cu:652-655): uses the dL_dtransMat values as stored in memory
    const float depth = precomp ? transMat_precomp[9 * (size_t)idx + 8] : splats[idx].Tw[2];
    const float g2 = wrote_back ? dT[0][2] : dT_u2, g5 = wrote_back ? dT[1][2] : dT_v2;
    dL_dmean2D[3 * (size_t)idx + 0] = g2 * depth * 0.5f * (float)W;
    dL_dmean2D[3 * (size_t)idx + 1] = g5 * depth * 0.5f * (float)H;
}

// ---- host launchers ---------------------------------------------------------------------------------------
int launch_preprocess_fwd(const IsrForwardArgs& a, cudaStream_t stream) {
    if (a.P == 0) return ISR_OK;
    GeomLayout gl(a.P);
    char* g = static_cast<char*>(a.geom);
    const int gx = (a.W + TILE - 1) / TILE, gy = (a.H + TILE - 1) / TILE;
    Camera cam{a.viewmatrix, a.projmatrix, a.campos};
    const size_t sh_smem = (a.shs != nullptr && a.sh_coeffs == 16 && a.colors_precomp == nullptr) ? 256 * 13 * 16 : 0;
    // ISR_K1_BULK=1: stage the SH rows with one cp.async.bulk per CTA instead of 12 cp.async per thread.  Measured
    // (profiles/r2_k1_bulk.md): 6.8% fewer warp instructions but 203 -> 246 us -- the copy lands unpadded (4-way bank
    // conflicts on the float4 reads) and every thread waits for the whole block's 48 KB.  Off by default.
    static const bool bulk = [] { const char* e = getenv("ISR_K1_BULK"); return e && e[0] == '1'; }();
    const bool aligned = ((uintptr_t)a.shs & 15) == 0;
    auto kern = (a.flags & ISR_FLAG_SPEC_ARITH)
                    ? (bulk && aligned ? preprocess_fwd_kernel<false, true> : preprocess_fwd_kernel<false, false>)
                    : (bulk && aligned ? preprocess_fwd_kernel<true, true> : preprocess_fwd_kernel<true, false>);
    ISR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 13 * 16));
    kern<<<(a.P + 255) / 256, 256, sh_smem, stream>>>(
        a.P, a.sh_degree, a.sh_coeffs, a.means3D, reinterpret_cast<const float2*>(a.scales), a.scale_modifier,
        reinterpret_cast<const float4*>(a.rotations), a.opacities, a.shs, a.transMat_precomp, a.colors_precomp, cam,
        a.W, a.H, gx, gy, a.radii, reinterpret_cast<Splat*>(g + gl.splat), reinterpret_cast<float4*>(g + gl.cull),
        reinterpret_cast<float4*>(g + gl.cullq), reinterpret_cast<float4*>(g + gl.rgb),
        reinterpret_cast<float*>(g + gl.depth), reinterpret_cast<uint32_t*>(g + gl.depth_key),
        reinterpret_cast<uint32_t*>(g + gl.tiles), reinterpret_cast<uint8_t*>(g + gl.clamped),
        reinterpret_cast<uint4*>(g + gl.tfoot), reinterpret_cast<uint32_t*>(g + gl.tcount)); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_preprocess_bwd(const IsrBackwardArgs& a, cudaStream_t stream) {
    if (a.P == 0) return ISR_OK;
    GeomLayout gl(a.P);
    const char* g = static_cast<const char*>(a.geom);
    int W = a.W, H = a.H;
    if (a.flags & ISR_FLAG_BWD_WH_QUIRK) {
        // backward.cu:633-634 with focal = W / (2 tan) from rasterizer_impl.cu:399-400, all in fp32
        volatile float focal_y = (float)a.H / (2.0f * a.tan_fovy);
        volatile float focal_x = (float)a.W / (2.0f * a.tan_fovx);
        volatile float fw = focal_x * a.tan_fovx;
        volatile float fh = focal_y * a.tan_fovy;
        volatile float fw2 = fw * 2.0f, fh2 = fh * 2.0f;
        W = (int)fw2;
        H = (int)fh2;
    }
    Camera cam{a.viewmatrix, a.projmatrix, a.campos};
    auto kern = (a.flags & ISR_FLAG_SPEC_ARITH) ? preprocess_bwd_kernel<false> : preprocess_bwd_kernel<true>;
    kern<<<(a.P + 255) / 256, 256, 0, stream>>>(
        a.P, a.sh_degree, a.sh_coeffs, a.means3D, a.radii, a.shs, reinterpret_cast<const uint8_t*>(g + gl.clamped),
        reinterpret_cast<const float2*>(a.scales), reinterpret_cast<const float4*>(a.rotations), a.transMat_precomp,
        reinterpret_cast<const Splat*>(g + gl.splat), cam, W, H, a.dL_dmeans2D, a.dL_dnormal, a.dL_dtransMat,
        a.dL_dcolors, a.dL_dsh, a.dL_dmeans3D, a.dL_dscales, a.dL_drotations); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

int launch_mark_visible(int P, const float* means3D, const float* view, const float* proj, uint8_t* present,
                        cudaStream_t stream) {
    if (P == 0) return ISR_OK;
    Camera cam{view, proj, nullptr};
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, cam, present); note_launch();
    ISR_CUDA_TRY(cudaGetLastError());
    return ISR_OK;
}

}  // namespace isr
