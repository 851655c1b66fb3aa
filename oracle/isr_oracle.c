/*
 * isr_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C, fp32 restatement of the reference 2D-Gaussian (surfel) rasterizer
 * zju3dv/InstaScene  submodules/diff-surfel-rasterization  (abbrev. DSR/), i.e. of
 *   DSR/cuda_rasterizer/forward.cu, backward.cu, rasterizer_impl.cu, auxiliary.h
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (instascene_b200/libisr.so) never links it.
 *
 * Arithmetic contract ("the spec").  The reference evaluates everything in fp32 with
 * nvcc's default FMA contraction, IEEE div+sqrt and CUDA's expf / rsqrtf.  This oracle
 * follows the operation sequence nvcc 12.9 emits for the reference (read from the SASS of
 * the unmodified build; noted inline as "[sass]"): same contraction, same operand order,
 * IEEE division where the reference divides.  The two MUFU-based functions, expf
 * (forward.cu:385) and rsqrtf (auxiliary.h:221), cannot be reproduced on a CPU and are
 * replaced by IEEE-only stand-ins: a Cody-Waite + degree-7 polynomial exp (orc_exp_neg)
 * and 1/sqrt.  The CUDA product has the same two stand-ins behind ISR_FLAG_SPEC_ARITH;
 * with that flag forward outputs -- including every thresholded integer (radii, tile
 * counts, sort order, n_contrib, pair list) -- are BIT-EXACT between this oracle and the
 * GPU.  Without the flag (the product default) the GPU evaluates expf / rsqrtf like the
 * reference and its forward is bit-identical to the unmodified reference CUDA build
 * (tests/test_reference_scale_gpu.py); this oracle then differs from both by the few ulp
 * of the two stand-ins, which is measured on B200 by tests/golden/make_golden.py and
 * pinned by the .npz files in tests/golden (integers exact, floats 1e-4).
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: the compiler must not fuse or
 * re-associate anything on its own).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BLOCK_X 16
#define BLOCK_Y 16
#define BLOCK_SIZE 256

/* DSR/cuda_rasterizer/auxiliary.h:38-41 */
static const float near_n = 0.2f;
static const float far_n = 100.0f;
static const float FilterSize = 0.707106f;
static const float FilterInvSquare = 2.0f;
/* DSR/cuda_rasterizer/auxiliary.h:44-61 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

#define ORC_FLAG_BWD_WH_QUIRK 1 /* Q5: backward recomputes W,H = int(focal*tan*2) (backward.cu:633-634) */

static inline float rcp_rn(float x) { return 1.0f / x; }
static inline float fminf_(float a, float b) { return a < b ? a : b; } /* args never NaN where used */
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }

/* float -> int32 with CUDA cvt.rzi.s32.f32 semantics (NaN -> 0, saturating). */
static inline int f2i_rz(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* exp(x) for x <= 0 (the blend only ever needs x in about [-5.6, 0]; valid to -80).
 * n = rint(x*log2e) via the 1.5*2^23 trick, Cody-Waite reduction, degree-7 Taylor/Horner,
 * scale by 2^n through the exponent bits.  |rel err| < 2 ulp.  x < -80 returns exactly 0
 * (the reference's expf underflows to a value whose alpha is < 1/255 for any sane opacity). */
static inline float orc_exp_neg(float x) {
    if (x < -80.0f) return 0.0f;
    const float LOG2E = 1.4426950408889634f, MAGIC = 12582912.0f;
    const float LN2_HI = 0.693145751953125f, LN2_LO = 1.42860682030941723212e-6f;
    float t = x * LOG2E;
    float tm = t + MAGIC;
    float n = tm - MAGIC;
    float r = fmaf(n, -LN2_HI, x);
    r = fmaf(n, -LN2_LO, r);
    float p = 1.984126984e-4f;          /* 1/5040 */
    p = fmaf(p, r, 1.388888889e-3f);    /* 1/720 */
    p = fmaf(p, r, 8.333333333e-3f);    /* 1/120 */
    p = fmaf(p, r, 4.166666667e-2f);    /* 1/24 */
    p = fmaf(p, r, 1.666666667e-1f);    /* 1/6 */
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    uint32_t bits;
    memcpy(&bits, &tm, 4);
    uint32_t sb = (bits << 23) + 0x3f800000u;
    float scale;
    memcpy(&scale, &sb, 4);
    return p * scale;
}

/* ------------------------------------------------------------------------------------------
 * getRect  (DSR/cuda_rasterizer/auxiliary.h:68-78)
 * ---------------------------------------------------------------------------------------- */
static inline void get_rect(float px, float py, int max_radius, int gx, int gy, int* mnx, int* mny,
                            int* mxx, int* mxy) {
    float r = (float)max_radius;
    *mnx = imin(gx, imax(0, f2i_rz((px - r) * 0.0625f)));
    *mny = imin(gy, imax(0, f2i_rz((py - r) * 0.0625f)));
    /* [sass] (p + r + BLOCK - 1) is evaluated left to right in fp32: ((p + r) + 16) - 1 */
    *mxx = imin(gx, imax(0, f2i_rz((((px + r) + 16.0f) - 1.0f) * 0.0625f)));
    *mxy = imin(gy, imax(0, f2i_rz((((py + r) + 16.0f) - 1.0f) * 0.0625f)));
}

/* quat (w,x,y,z stored in columns 0..3) -> rotation columns R0,R1,R2
 * (DSR/cuda_rasterizer/auxiliary.h:214-236).  rsqrtf is replaced by 1/sqrt (spec). */
static inline void quat_to_rot(const float* q, float R0[3], float R1[3], float R2[3]) {
    /* [sass] sum = fma(q2,q2, fma(q1,q1, fma(q0,q0, q3*q3))) */
    float sum = fmaf(q[2], q[2], fmaf(q[1], q[1], fmaf(q[0], q[0], q[3] * q[3])));
    float s = rcp_rn(sqrtf(sum));
    float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
    /* [sass] y*y and z*z are shared (rounded) products: R0[0] adds them with a plain FADD */
    float yy = y * y, zz = z * z;
    float yy_zz = yy + zz, xx_zz = fmaf(x, x, zz), xx_yy = fmaf(x, x, yy);
    float xy_p = fmaf(x, y, w * z), xy_m = fmaf(x, y, -(w * z));
    float xz_p = fmaf(x, z, w * y), xz_m = fmaf(x, z, -(w * y));
    float yz_p = fmaf(y, z, w * x), yz_m = fmaf(y, z, -(w * x));
    R0[0] = 1.0f - (yy_zz + yy_zz); R0[1] = xy_p + xy_p;          R0[2] = xz_m + xz_m;
    R1[0] = xy_m + xy_m;          R1[1] = 1.0f - (xx_zz + xx_zz); R1[2] = yz_p + yz_p;
    R2[0] = xz_p + xz_p;          R2[1] = yz_m + yz_m;          R2[2] = 1.0f - (xx_yy + xx_yy);
}

/* a*x + b*y + c*z evaluated as fma(c,z, fma(a,x, b*y)) -- nvcc's contraction of a left-to-right
 * three-term dot product. */
static inline float dot3c(float a, float x, float b, float y, float c, float z) {
    return fmaf(c, z, fmaf(a, x, b * y));
}

/* compute_transmat (DSR/cuda_rasterizer/forward.cu:75-115).  T rows Tu,Tv,Tw -> T[0..8]. */
static void compute_transmat(const float* p, float sx, float sy, const float* q, const float* proj,
                             const float* view, int W, int H, float T[9], float normal[3]) {
    float R0[3], R1[3], R2[3];
    quat_to_rot(q, R0, R1, R2);
    float L0[3] = {R0[0] * sx, R0[1] * sx, R0[2] * sx};
    float L1[3] = {R1[0] * sy, R1[1] * sy, R1[2] * sy};
    /* X = transpose(splat2world) * world2ndc ; X[c][r], c = ndc column (0..3), r = (u,v,1) */
    float X[4][3];
    for (int c = 0; c < 4; c++) {
        X[c][0] = dot3c(L0[0], proj[c], L0[1], proj[4 + c], L0[2], proj[8 + c]);
        X[c][1] = dot3c(L1[0], proj[c], L1[1], proj[4 + c], L1[2], proj[8 + c]);
        X[c][2] = dot3c(p[0], proj[c], p[1], proj[4 + c], p[2], proj[8 + c]) + proj[12 + c];
    }
    const float hw = (float)W * 0.5f, hw1 = (float)(W - 1) * 0.5f;
    const float hh = (float)H * 0.5f, hh1 = (float)(H - 1) * 0.5f;
    for (int r = 0; r < 3; r++) {
        T[0 + r] = fmaf(X[3][r], hw1, X[0][r] * hw);
        T[3 + r] = fmaf(X[3][r], hh1, X[1][r] * hh);
        T[6 + r] = X[3][r];
    }
    /* normal = transformVec4x3(L[2], view)  (auxiliary.h:101-109) */
    normal[0] = dot3c(view[0], R2[0], view[4], R2[1], view[8], R2[2]);
    normal[1] = dot3c(view[1], R2[0], view[5], R2[1], view[9], R2[2]);
    normal[2] = dot3c(view[2], R2[0], view[6], R2[1], view[10], R2[2]);
}

/* compute_aabb (DSR/cuda_rasterizer/forward.cu:119-145), cutoff = 3.  FMA pattern = [sass]. */
static int compute_aabb(const float T[9], float* cx, float* cy, float* ex, float* ey) {
    const float *Tu = T, *Tv = T + 3, *Tw = T + 6;
    float d = fmaf(-Tw[2], Tw[2], fmaf(Tw[0] * Tw[0], 9.0f, (Tw[1] * Tw[1]) * 9.0f));
    if (d == 0.0f) return 0;
    float inv_d = rcp_rn(d);
    float f9 = inv_d * 9.0f;
    /* [sass] every dot(f, a*b): fma(a.z*b.z, -+inv_d, fma(f9, a.x*b.x, f9*(a.y*b.y))) -- the FMA carries the x term */
    float px = fmaf(Tu[2] * Tw[2], -inv_d, fmaf(f9, Tu[0] * Tw[0], f9 * (Tu[1] * Tw[1])));
    float py = fmaf(Tv[2] * Tw[2], -inv_d, fmaf(f9, Tv[0] * Tw[0], f9 * (Tv[1] * Tw[1])));
    float nx = fmaf(Tu[2] * Tu[2], inv_d, -fmaf(f9, Tu[0] * Tu[0], f9 * (Tu[1] * Tu[1])));
    float ny = fmaf(Tv[2] * Tv[2], inv_d, -fmaf(f9, Tv[0] * Tv[0], f9 * (Tv[1] * Tv[1])));
    float h0x = fmaf(px, px, nx), h0y = fmaf(py, py, ny);
    *cx = px; *cy = py;
    *ex = sqrtf(fmaxf_(1e-4f, h0x));
    *ey = sqrtf(fmaxf_(1e-4f, h0y));
    return 1;
}

/* computeColorFromSH forward (DSR/cuda_rasterizer/forward.cu:20-71). */
static void sh_to_rgb(int deg, const float* pos, const float* campos, const float* sh /*[M][3]*/,
                      float rgb[3], uint8_t clamped[3]) {
    float dx = pos[0] - campos[0], dy = pos[1] - campos[1], dz = pos[2] - campos[2];
    float len = sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
    float x = dx / len, y = dy / len, z = dz / len;
    float res[3];
    for (int c = 0; c < 3; c++) res[c] = SH_C0 * sh[c];
    if (deg > 0) {
        float a1 = SH_C1 * y, a2 = SH_C1 * z, a3 = SH_C1 * x;
        for (int c = 0; c < 3; c++) {
            float r = res[c];
            r = fmaf(-a1, sh[3 + c], r);
            r = fmaf(a2, sh[6 + c], r);
            r = fmaf(-a3, sh[9 + c], r);
            res[c] = r;
        }
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            float b4 = SH_C2[0] * xy, b5 = SH_C2[1] * yz, b6 = SH_C2[2] * ((2.0f * zz - xx) - yy);
            float b7 = SH_C2[3] * xz, b8 = SH_C2[4] * (xx - yy);
            for (int c = 0; c < 3; c++) {
                float r = res[c];
                r = fmaf(b4, sh[12 + c], r);
                r = fmaf(b5, sh[15 + c], r);
                r = fmaf(b6, sh[18 + c], r);
                r = fmaf(b7, sh[21 + c], r);
                r = fmaf(b8, sh[24 + c], r);
                res[c] = r;
            }
            if (deg > 2) {
                /* [sass] 3xx-yy = fma(xx,3,-yy); 4zz-xx-yy = fma(zz,4,-xx) - yy (shared by c11, c13);
                 * 2zz-3xx-3yy = fma(yy,-3, fma(xx,-3, zz+zz)); xx-3yy = fma(yy,-3,xx) */
                float t4 = fmaf(zz, 4.0f, -xx) - yy;
                float c9 = (SH_C3[0] * y) * fmaf(xx, 3.0f, -yy);
                float c10 = (SH_C3[1] * xy) * z;
                float c11 = (SH_C3[2] * y) * t4;
                float c12 = (SH_C3[3] * z) * fmaf(yy, -3.0f, fmaf(xx, -3.0f, zz + zz));
                float c13 = (SH_C3[4] * x) * t4;
                float c14 = (SH_C3[5] * z) * (xx - yy);
                float c15 = (SH_C3[6] * x) * fmaf(yy, -3.0f, xx);
                for (int c = 0; c < 3; c++) {
                    float r = res[c];
                    r = fmaf(c9, sh[27 + c], r);
                    r = fmaf(c10, sh[30 + c], r);
                    r = fmaf(c11, sh[33 + c], r);
                    r = fmaf(c12, sh[36 + c], r);
                    r = fmaf(c13, sh[39 + c], r);
                    r = fmaf(c14, sh[42 + c], r);
                    r = fmaf(c15, sh[45 + c], r);
                    res[c] = r;
                }
            }
        }
    }
    for (int c = 0; c < 3; c++) {
        float r = res[c] + 0.5f;
        clamped[c] = (r < 0.0f);
        rgb[c] = fmaxf_(r, 0.0f);
    }
}

/* ------------------------------------------------------------------------------------------
 * K1 preprocess forward  (DSR/cuda_rasterizer/forward.cu:148-251, auxiliary.h:186-211)
 * All output arrays have P rows; rows of culled Gaussians are left untouched except
 * radii/tiles_touched = 0 (exactly as the reference, which leaves stale memory there).
 * ---------------------------------------------------------------------------------------- */
void orc_preprocess_forward(int P, int D, int M, const float* means3D, const float* scales,
                            float scale_modifier, const float* rotations, const float* opacities,
                            const float* shs, const float* transMat_precomp,
                            const float* colors_precomp, const float* view, const float* proj,
                            const float* campos, int W, int H, int* radii, float* means2D,
                            float* depths, float* transMats, float* rgb, float* normal_opacity,
                            uint32_t* tiles_touched, uint8_t* clamped) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        radii[idx] = 0;
        tiles_touched[idx] = 0;
        const float* p = means3D + 3 * (size_t)idx;
        /* p_view.z : transformPoint4x3 (auxiliary.h:80-88) */
        float pvx = dot3c(view[0], p[0], view[4], p[1], view[8], p[2]) + view[12];
        float pvy = dot3c(view[1], p[0], view[5], p[1], view[9], p[2]) + view[13];
        float pvz = dot3c(view[2], p[0], view[6], p[1], view[10], p[2]) + view[14];
        if (pvz <= 0.2f) continue;
        float T[9], normal[3];
        if (transMat_precomp == NULL) {
            compute_transmat(p, scale_modifier * scales[2 * (size_t)idx],
                             scale_modifier * scales[2 * (size_t)idx + 1], rotations + 4 * (size_t)idx,
                             proj, view, W, H, T, normal);
            memcpy(transMats + 9 * (size_t)idx, T, sizeof(T));
        } else {
            memcpy(T, transMat_precomp + 9 * (size_t)idx, sizeof(T));
            normal[0] = 0.0f; normal[1] = 0.0f; normal[2] = 1.0f;
        }
        /* DUAL_VISIABLE (forward.cu:209-214) */
        float cosv = -fmaf(pvz, normal[2], fmaf(pvx, normal[0], pvy * normal[1]));
        if (cosv == 0.0f) continue;
        float mult = cosv > 0.0f ? 1.0f : -1.0f;
        normal[0] *= mult; normal[1] *= mult; normal[2] *= mult;
        float cx, cy, ex, ey;
        if (!compute_aabb(T, &cx, &cy, &ex, &ey)) continue;
        float radius = ceilf(fmaxf_(fmaxf_(ex, ey), 3.0f * FilterSize));
        int mnx, mny, mxx, mxy;
        get_rect(cx, cy, f2i_rz(radius), gx, gy, &mnx, &mny, &mxx, &mxy);
        if ((mxx - mnx) * (mxy - mny) == 0) continue;
        if (colors_precomp == NULL) {
            sh_to_rgb(D, p, campos, shs + (size_t)idx * M * 3, rgb + 3 * (size_t)idx,
                      clamped + 3 * (size_t)idx);
        }
        depths[idx] = pvz;
        radii[idx] = f2i_rz(radius);
        means2D[2 * (size_t)idx] = cx;
        means2D[2 * (size_t)idx + 1] = cy;
        normal_opacity[4 * (size_t)idx + 0] = normal[0];
        normal_opacity[4 * (size_t)idx + 1] = normal[1];
        normal_opacity[4 * (size_t)idx + 2] = normal[2];
        normal_opacity[4 * (size_t)idx + 3] = opacities[idx];
        tiles_touched[idx] = (uint32_t)((mxy - mny) * (mxx - mnx));
    }
}

/* markVisible / checkFrustum (DSR/cuda_rasterizer/rasterizer_impl.cu:54-66) */
void orc_mark_visible(int P, const float* means3D, const float* view, uint8_t* present) {
    for (int idx = 0; idx < P; idx++) {
        const float* p = means3D + 3 * (size_t)idx;
        float pvz = dot3c(view[2], p[0], view[6], p[1], view[10], p[2]) + view[14];
        present[idx] = pvz > 0.2f;
    }
}

/* ------------------------------------------------------------------------------------------
 * K2-K5 binning (DSR/cuda_rasterizer/rasterizer_impl.cu:70-111, 116-138, 283-324):
 * inclusive scan, key emission in Gaussian-id order, STABLE ascending sort of
 * (tile<<32 | depth_bits), per-tile [start,end).  Returns R.  point_list/keys sized >= R.
 * ---------------------------------------------------------------------------------------- */
int64_t orc_bin_count(int P, const uint32_t* tiles_touched, uint32_t* point_offsets) {
    uint32_t acc = 0;
    for (int i = 0; i < P; i++) { acc += tiles_touched[i]; point_offsets[i] = acc; }
    return (int64_t)acc;
}

static void radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* ktmp, uint32_t* vtmp, size_t n,
                             int bits) {
    for (int shift = 0; shift < bits; shift += 8) {
        size_t hist[257];
        memset(hist, 0, sizeof(hist));
        for (size_t i = 0; i < n; i++) hist[((keys[i] >> shift) & 0xff) + 1]++;
        for (int b = 0; b < 256; b++) hist[b + 1] += hist[b];
        for (size_t i = 0; i < n; i++) {
            size_t d = hist[(keys[i] >> shift) & 0xff]++;
            ktmp[d] = keys[i];
            vtmp[d] = vals[i];
        }
        memcpy(keys, ktmp, n * sizeof(uint64_t));
        memcpy(vals, vtmp, n * sizeof(uint32_t));
    }
}

void orc_bin(int P, int W, int H, const int* radii, const float* means2D, const float* depths,
             const uint32_t* point_offsets, int64_t R, uint64_t* keys_sorted, uint32_t* point_list,
             uint32_t* ranges /*[tiles][2]*/) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
    if (R == 0) return;
    for (int idx = 0; idx < P; idx++) {
        if (radii[idx] <= 0) continue;
        uint32_t off = idx == 0 ? 0 : point_offsets[idx - 1];
        int mnx, mny, mxx, mxy;
        get_rect(means2D[2 * (size_t)idx], means2D[2 * (size_t)idx + 1], radii[idx], gx, gy, &mnx, &mny,
                 &mxx, &mxy);
        uint32_t dbits;
        memcpy(&dbits, &depths[idx], 4);
        for (int y = mny; y < mxy; y++)
            for (int x = mnx; x < mxx; x++) {
                uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
                key = (key << 32) | dbits;
                keys_sorted[off] = key;
                point_list[off] = (uint32_t)idx;
                off++;
            }
    }
    uint64_t* ktmp = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)R);
    uint32_t* vtmp = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)R);
    radix_sort_pairs(keys_sorted, point_list, ktmp, vtmp, (size_t)R, 64);
    free(ktmp);
    free(vtmp);
    /* identifyTileRanges */
    for (int64_t i = 0; i < R; i++) {
        uint32_t cur = (uint32_t)(keys_sorted[i] >> 32);
        if (i == 0) ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(keys_sorted[i - 1] >> 32);
            if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
    }
}

/* ------------------------------------------------------------------------------------------
 * Per-(pixel, Gaussian) evaluation shared by forward and backward
 * (DSR/cuda_rasterizer/forward.cu:355-393 == backward.cu:293-325).
 * Returns 0 if this pair is skipped.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    float kx, ky, kz, lx, ly, lz, px, py, pz, sx, sy, rho3d, rho2d, ddx, ddy, depth, G, alpha;
    int use3d;
} PairEval;

static inline int eval_pair(float pixx, float pixy, const float* xy, const float* T, float opa,
                            PairEval* e) {
    const float *Tu = T, *Tv = T + 3, *Tw = T + 6;
    e->kx = fmaf(pixx, Tw[0], -Tu[0]); e->ky = fmaf(pixx, Tw[1], -Tu[1]); e->kz = fmaf(pixx, Tw[2], -Tu[2]);
    e->lx = fmaf(pixy, Tw[0], -Tv[0]); e->ly = fmaf(pixy, Tw[1], -Tv[1]); e->lz = fmaf(pixy, Tw[2], -Tv[2]);
    e->px = fmaf(e->ky, e->lz, -(e->kz * e->ly));
    e->py = fmaf(e->kz, e->lx, -(e->kx * e->lz));
    e->pz = fmaf(e->kx, e->ly, -(e->ky * e->lx));
    if (e->pz == 0.0f) return 0;
    /* [sass] s = p.xy / p.z is an IEEE division (div.rn.f32) for any non-zero p.z */
    e->sx = e->px / e->pz;
    e->sy = e->py / e->pz;
    e->rho3d = fmaf(e->sx, e->sx, e->sy * e->sy);
    e->ddx = xy[0] - pixx;
    e->ddy = xy[1] - pixy;
    /* [sass] FMUL d.x*d.x; FFMA d.y,d.y; FADD r+r */
    e->rho2d = FilterInvSquare * fmaf(e->ddy, e->ddy, e->ddx * e->ddx);
    e->use3d = e->rho3d <= e->rho2d;
    /* FMNMX: the non-NaN operand if one is NaN */
    float rho = (e->rho3d != e->rho3d) ? e->rho2d : ((e->rho2d != e->rho2d) ? e->rho3d : fminf_(e->rho3d, e->rho2d));
    /* [sass] fma(Tw.x, s.x, Tw.y*s.y) + Tw.z */
    e->depth = e->use3d ? fmaf(Tw[0], e->sx, Tw[1] * e->sy) + Tw[2] : Tw[2];
    if (e->depth < near_n) return 0;
    float power = rho * -0.5f;
    if (power > 0.0f) return 0;
    e->G = orc_exp_neg(power);
    e->alpha = fminf_(0.99f, opa * e->G);
    if (e->alpha < 1.0f / 255.0f) return 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * K6 blend forward (DSR/cuda_rasterizer/forward.cu:256-462).
 * out_color[3,H,W], out_others[7,H,W], out_extra[F,H,W], final_T[3,H,W], n_contrib[2,H,W],
 * pairs[cap][2] + *pair_count (#pairs; order = pixel-major then list order, compare as a set).
 * ---------------------------------------------------------------------------------------- */
void orc_blend_forward(int W, int H, int F, const uint32_t* ranges, const uint32_t* point_list,
                       const float* means2D, const float* colors, const float* transMats,
                       const float* extras, const float* normal_opacity, const float* bg,
                       float* final_T, uint32_t* n_contrib, float* out_color, float* out_others,
                       float* out_extra, int32_t* pairs, int64_t pair_cap, int64_t* pair_count,
                       int tile_stride) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const size_t HW = (size_t)H * W;
    const float c1 = far_n / (far_n - near_n);
    int64_t npairs = 0;
    if (tile_stride < 1) tile_stride = 1; /* >1: only every tile_stride-th tile (bounded CPU-baseline sample) */
#pragma omp parallel for schedule(dynamic, 4) collapse(2)
    for (int ty = 0; ty < gy; ty++)
        for (int tx = 0; tx < gx; tx++) {
            if ((ty * gx + tx) % tile_stride != 0) continue;
            const uint32_t r0 = ranges[2 * ((size_t)ty * gx + tx)], r1 = ranges[2 * ((size_t)ty * gx + tx) + 1];
            float* E = (float*)malloc(sizeof(float) * (size_t)(F > 0 ? F : 1));
            for (int ly = 0; ly < BLOCK_Y; ly++)
                for (int lx = 0; lx < BLOCK_X; lx++) {
                    const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
                    if (pxi >= W || pyi >= H) continue;
                    const size_t pix_id = (size_t)W * pyi + pxi;
                    const float pixx = (float)pxi, pixy = (float)pyi;
                    float T = 1.0f, C[3] = {0, 0, 0}, N[3] = {0, 0, 0};
                    float D = 0, M1 = 0, M2 = 0, distortion = 0, median_depth = 0;
                    uint32_t contributor = 0, last_contributor = 0, median_contributor = 0;
                    for (int ch = 0; ch < F; ch++) E[ch] = 0.0f;
                    for (uint32_t i = r0; i < r1; i++) {
                        contributor++;
                        const uint32_t g = point_list[i];
                        const float* no = normal_opacity + 4 * (size_t)g;
                        PairEval e;
                        if (!eval_pair(pixx, pixy, means2D + 2 * (size_t)g, transMats + 9 * (size_t)g, no[3], &e))
                            continue;
                        float test_T = T * (1.0f - e.alpha);
                        if (test_T < 0.0001f) break; /* done = true */
                        float w = e.alpha * T;
                        float A = 1.0f - T;
                        float m = c1 * (1.0f - near_n / e.depth); /* [sass] div.rn, FADD, FMUL */
                        float mm = m * m;
                        float dt = fmaf(-(m + m), M1, fmaf(mm, A, M2));
                        distortion = fmaf(dt, w, distortion);
                        D = fmaf(e.depth, w, D);
                        M1 = fmaf(m, w, M1);
                        M2 = fmaf(mm, w, M2);
                        if (T > 0.5f) { median_depth = e.depth; median_contributor = contributor; }
                        for (int ch = 0; ch < 3; ch++) N[ch] = fmaf(no[ch], w, N[ch]);
                        /* [sass] forward.cu:415 extras * alpha * T: fma(T, alpha*feature, E) */
                        for (int ch = 0; ch < F; ch++) E[ch] = fmaf(T, extras[(size_t)g * F + ch] * e.alpha, E[ch]);
                        for (int ch = 0; ch < 3; ch++) C[ch] = fmaf(colors[3 * (size_t)g + ch], w, C[ch]);
                        if (w >= 0.1f) { /* reference: (double)w > 0.1  <=>  w >= 0.1f  (forward.cu:422) */
                            int64_t slot;
#pragma omp atomic capture
                            slot = npairs++;
                            if (pairs && slot < pair_cap) {
                                pairs[2 * slot] = (int32_t)g;
                                pairs[2 * slot + 1] = (int32_t)pix_id;
                            }
                        }
                        T = test_T;
                        last_contributor = contributor;
                    }
                    final_T[pix_id] = T;
                    final_T[pix_id + HW] = M1;
                    final_T[pix_id + 2 * HW] = M2;
                    n_contrib[pix_id] = last_contributor;
                    n_contrib[pix_id + HW] = median_contributor; /* Q3: float -1 -> u32 saturates to 0 */
                    for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix_id] = fmaf(T, bg[ch], C[ch]);
                    out_others[pix_id + 0 * HW] = D;
                    out_others[pix_id + 1 * HW] = 1.0f - T;
                    for (int ch = 0; ch < 3; ch++) out_others[pix_id + (2 + ch) * HW] = N[ch];
                    out_others[pix_id + 5 * HW] = median_depth;
                    out_others[pix_id + 6 * HW] = distortion;
                    for (int ch = 0; ch < F; ch++) out_extra[ch * HW + pix_id] = E[ch];
                }
            free(E);
        }
    *pair_count = npairs;
}

/* ------------------------------------------------------------------------------------------
 * K7 blend backward (DSR/cuda_rasterizer/backward.cu:143-466).  Gradient buffers must be
 * zero-initialised by the caller; accumulation order = tile-major, pixel-major, back-to-front
 * (the reference's order is non-deterministic: float atomics).
 * ---------------------------------------------------------------------------------------- */
void orc_blend_backward(int W, int H, int F, const uint32_t* ranges, const uint32_t* point_list,
                        const float* bg, const float* means2D, const float* normal_opacity,
                        const float* transMats, const float* colors, const float* extras,
                        const float* final_T, const uint32_t* n_contrib, const float* dL_dpixels,
                        const float* dL_dothers, const float* dL_dpixel_extras, float* dL_dtransMat,
                        float* dL_dmean2D /*[P][3]*/, float* dL_dnormal3D, float* dL_dopacity,
                        float* dL_dcolors, float* dL_dextras, int tile_stride,
                        const uint8_t* pixel_mask /* [H*W] or NULL: only pixels with a non-zero entry are walked (the
                                                     caller guarantees every cotangent of the others is zero) */) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const size_t HW = (size_t)H * W;
    const float c1 = far_n / (far_n - near_n);
    const float c3 = (far_n * near_n) / (far_n - near_n);
    const int Fa = F > 0 ? F : 1;
    if (tile_stride < 1) tile_stride = 1;
    /* Tiles run in parallel; per-Gaussian sums use atomic adds (like the reference's float atomics the order is
     * then unspecified).  With orc_set_num_threads(1) the order is tile-major, pixel-major, back-to-front. */
#define ACC(lhs, val) do { const float v_ = (val); _Pragma("omp atomic") lhs += v_; } while (0)
#pragma omp parallel for schedule(dynamic, 4) collapse(2)
    for (int ty = 0; ty < gy; ty++)
        for (int tx = 0; tx < gx; tx++) {
            if ((ty * gx + tx) % tile_stride != 0) continue;
            float* accum_ree = (float*)malloc(sizeof(float) * 3 * Fa);
            float* last_extra = accum_ree + Fa;
            float* dLdE = accum_ree + 2 * Fa;
            const uint32_t r0 = ranges[2 * ((size_t)ty * gx + tx)], r1 = ranges[2 * ((size_t)ty * gx + tx) + 1];
            for (int ly = 0; ly < BLOCK_Y; ly++)
                for (int lx = 0; lx < BLOCK_X; lx++) {
                    const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
                    if (pxi >= W || pyi >= H) continue;
                    const size_t pix_id = (size_t)W * pyi + pxi;
                    if (pixel_mask && !pixel_mask[pix_id]) continue;
                    const float pixx = (float)pxi, pixy = (float)pyi;
                    const float T_final = final_T[pix_id];
                    float T = T_final;
                    const uint32_t last_contributor = n_contrib[pix_id];
                    const uint32_t median_contributor = n_contrib[pix_id + HW];
                    float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, dL_dpixel[3];
                    for (int ch = 0; ch < 3; ch++) dL_dpixel[ch] = dL_dpixels[ch * HW + pix_id];
                    const float dL_ddepth = dL_dothers[0 * HW + pix_id];
                    const float dL_daccum = dL_dothers[1 * HW + pix_id];
                    const float dL_dreg = dL_dothers[6 * HW + pix_id];
                    float dL_dnormal2D[3];
                    for (int ch = 0; ch < 3; ch++) dL_dnormal2D[ch] = dL_dothers[(2 + ch) * HW + pix_id];
                    const float dL_dmedian_depth = dL_dothers[5 * HW + pix_id];
                    for (int ch = 0; ch < F; ch++) {
                        dLdE[ch] = dL_dpixel_extras[ch * HW + pix_id];
                        accum_ree[ch] = 0.0f;
                        last_extra[ch] = 0.0f;
                    }
                    float last_depth = 0, last_normal[3] = {0, 0, 0};
                    float accum_depth_rec = 0, accum_alpha_rec = 0, accum_normal_rec[3] = {0, 0, 0};
                    const float final_D = final_T[pix_id + HW], final_D2 = final_T[pix_id + 2 * HW];
                    const float final_A = 1.0f - T_final;
                    float last_dL_dT = 0, last_alpha = 0;
                    float bg_dot_dpixel = fmaf(bg[2], dL_dpixel[2], fmaf(bg[1], dL_dpixel[1], bg[0] * dL_dpixel[0]));
                    /* contributor index c (1-based in forward) runs last_contributor .. 1 */
                    for (uint32_t c = last_contributor; c >= 1; c--) {
                        const uint32_t contributor = c - 1; /* == reference's `contributor` after decrement */
                        const uint32_t g = point_list[r0 + contributor];
                        (void)r1;
                        const float* no = normal_opacity + 4 * (size_t)g;
                        const float* Tm = transMats + 9 * (size_t)g;
                        PairEval e;
                        if (!eval_pair(pixx, pixy, means2D + 2 * (size_t)g, Tm, no[3], &e)) continue;
                        const float alpha = e.alpha, G = e.G, c_d = e.depth;
                        const float ra = rcp_rn(1.0f - alpha);
                        T = T * ra;
                        const float w = alpha * T;
                        float dL_dalpha = 0.0f;
                        const float one_m_la = 1.0f - last_alpha;
                        for (int ch = 0; ch < 3; ch++) {
                            const float col = colors[3 * (size_t)g + ch];
                            accum_rec[ch] = fmaf(last_alpha, last_color[ch], one_m_la * accum_rec[ch]);
                            last_color[ch] = col;
                            dL_dalpha = fmaf(col - accum_rec[ch], dL_dpixel[ch], dL_dalpha);
                            ACC(dL_dcolors[3 * (size_t)g + ch], w * dL_dpixel[ch]);
                        }
                        float dL_dz = 0.0f;
                        const float rcd = rcp_rn(c_d);
                        const float m_d = c1 * (1.0f - near_n * rcd);
                        const float dmd_dd = (c3 * rcd) * rcd;
                        if (contributor == median_contributor - 1u) dL_dz += dL_dmedian_depth;
                        const float mm = m_d * m_d;
                        const float dL_dweight = fmaf(-(m_d + m_d), final_D, fmaf(mm, final_A, final_D2)) * dL_dreg;
                        dL_dalpha += dL_dweight - last_dL_dT;
                        last_dL_dT = fmaf(dL_dweight, alpha, (1.0f - alpha) * last_dL_dT);
                        const float dL_dmd = ((w + w) * fmaf(m_d, final_A, -final_D)) * dL_dreg;
                        dL_dz = fmaf(dL_dmd, dmd_dd, dL_dz);
                        accum_depth_rec = fmaf(last_alpha, last_depth, one_m_la * accum_depth_rec);
                        last_depth = c_d;
                        dL_dalpha = fmaf(c_d - accum_depth_rec, dL_ddepth, dL_dalpha);
                        accum_alpha_rec = last_alpha + one_m_la * accum_alpha_rec;
                        dL_dalpha = fmaf(1.0f - accum_alpha_rec, dL_daccum, dL_dalpha);
                        for (int ch = 0; ch < 3; ch++) {
                            accum_normal_rec[ch] = fmaf(last_alpha, last_normal[ch], one_m_la * accum_normal_rec[ch]);
                            last_normal[ch] = no[ch];
                            dL_dalpha = fmaf(no[ch] - accum_normal_rec[ch], dL_dnormal2D[ch], dL_dalpha);
                            ACC(dL_dnormal3D[3 * (size_t)g + ch], w * dL_dnormal2D[ch]);
                        }
                        for (int ch = 0; ch < F; ch++) {
                            const float ex = extras[(size_t)g * F + ch];
                            accum_ree[ch] = fmaf(last_alpha, last_extra[ch], one_m_la * accum_ree[ch]);
                            last_extra[ch] = ex;
                            dL_dalpha = fmaf(ex - accum_ree[ch], dLdE[ch], dL_dalpha);
                            ACC(dL_dextras[(size_t)g * F + ch], w * dLdE[ch]);
                        }
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha = fmaf((-T_final) * ra, bg_dot_dpixel, dL_dalpha);
                        const float dL_dG = no[3] * dL_dalpha;
                        dL_dz = fmaf(w, dL_ddepth, dL_dz);
                        float* dT = dL_dtransMat + 9 * (size_t)g;
                        if (e.use3d) {
                            const float nGd = dL_dG * (-G);
                            const float dsx = fmaf(nGd, e.sx, dL_dz * Tm[6]);
                            const float dsy = fmaf(nGd, e.sy, dL_dz * Tm[7]);
                            const float rpz = rcp_rn(e.pz);
                            const float dsx_pz = dsx * rpz, dsy_pz = dsy * rpz;
                            const float dpx = dsx_pz, dpy = dsy_pz, dpz = -fmaf(dsx_pz, e.sx, dsy_pz * e.sy);
                            /* dL_dk = cross(l, dL_dp); dL_dl = cross(dL_dp, k) */
                            const float dkx = fmaf(e.ly, dpz, -(e.lz * dpy));
                            const float dky = fmaf(e.lz, dpx, -(e.lx * dpz));
                            const float dkz = fmaf(e.lx, dpy, -(e.ly * dpx));
                            const float dlx = fmaf(dpy, e.kz, -(dpz * e.ky));
                            const float dly = fmaf(dpz, e.kx, -(dpx * e.kz));
                            const float dlz = fmaf(dpx, e.ky, -(dpy * e.kx));
                            ACC(dT[0], -dkx); ACC(dT[1], -dky); ACC(dT[2], -dkz);
                            ACC(dT[3], -dlx); ACC(dT[4], -dly); ACC(dT[5], -dlz);
                            ACC(dT[6], fmaf(dL_dz, e.sx, fmaf(pixx, dkx, pixy * dlx)));
                            ACC(dT[7], fmaf(dL_dz, e.sy, fmaf(pixx, dky, pixy * dly)));
                            ACC(dT[8], fmaf(pixx, dkz, pixy * dlz) + dL_dz);
                        } else {
                            const float dG_ddelx = (-G * FilterInvSquare) * e.ddx;
                            const float dG_ddely = (-G * FilterInvSquare) * e.ddy;
                            ACC(dL_dmean2D[3 * (size_t)g + 0], dL_dG * dG_ddelx);
                            ACC(dL_dmean2D[3 * (size_t)g + 1], dL_dG * dG_ddely);
                            ACC(dT[8], dL_dz);
                        }
                        ACC(dL_dopacity[g], G * dL_dalpha);
                    }
                }
            free(accum_ree);
        }
#undef ACC
}

/* ------------------------------------------------------------------------------------------
 * K8 preprocess backward (DSR/cuda_rasterizer/backward.cu:469-656, 20-139;
 * auxiliary.h:121-154, 239-283).  Float-only results (tolerance-compared), so plain C
 * expressions are used; -ffp-contract=off keeps them deterministic.
 * ---------------------------------------------------------------------------------------- */
static void sh_backward(int idx, int deg, int M, const float* means, const float* campos,
                        const float* shs, const uint8_t* clamped, const float* dL_dcolor,
                        float* dL_dmeans, float* dL_dshs) {
    const float* pos = means + 3 * (size_t)idx;
    float dox = pos[0] - campos[0], doy = pos[1] - campos[1], doz = pos[2] - campos[2];
    float len = sqrtf(dox * dox + doy * doy + doz * doz);
    float x = dox / len, y = doy / len, z = doz / len;
    const float* sh = shs + (size_t)idx * M * 3;
    float dRGB[3];
    for (int c = 0; c < 3; c++) dRGB[c] = dL_dcolor[3 * (size_t)idx + c] * (clamped[3 * (size_t)idx + c] ? 0.0f : 1.0f);
    float dx[3] = {0, 0, 0}, dy[3] = {0, 0, 0}, dz[3] = {0, 0, 0};
    float* dsh = dL_dshs + (size_t)idx * M * 3;
#define SH(i, c) sh[3 * (i) + (c)]
#define DSH(i, v) for (int c = 0; c < 3; c++) dsh[3 * (i) + c] = (v) * dRGB[c]
    DSH(0, SH_C0);
    if (deg > 0) {
        DSH(1, -SH_C1 * y); DSH(2, SH_C1 * z); DSH(3, -SH_C1 * x);
        for (int c = 0; c < 3; c++) { dx[c] = -SH_C1 * SH(3, c); dy[c] = -SH_C1 * SH(1, c); dz[c] = SH_C1 * SH(2, c); }
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            DSH(4, SH_C2[0] * xy); DSH(5, SH_C2[1] * yz); DSH(6, SH_C2[2] * (2.f * zz - xx - yy));
            DSH(7, SH_C2[3] * xz); DSH(8, SH_C2[4] * (xx - yy));
            for (int c = 0; c < 3; c++) {
                dx[c] += SH_C2[0] * y * SH(4, c) + SH_C2[2] * 2.f * -x * SH(6, c) + SH_C2[3] * z * SH(7, c) + SH_C2[4] * 2.f * x * SH(8, c);
                dy[c] += SH_C2[0] * x * SH(4, c) + SH_C2[1] * z * SH(5, c) + SH_C2[2] * 2.f * -y * SH(6, c) + SH_C2[4] * 2.f * -y * SH(8, c);
                dz[c] += SH_C2[1] * y * SH(5, c) + SH_C2[2] * 2.f * 2.f * z * SH(6, c) + SH_C2[3] * x * SH(7, c);
            }
            if (deg > 2) {
                DSH(9, SH_C3[0] * y * (3.f * xx - yy)); DSH(10, SH_C3[1] * xy * z);
                DSH(11, SH_C3[2] * y * (4.f * zz - xx - yy));
                DSH(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                DSH(13, SH_C3[4] * x * (4.f * zz - xx - yy)); DSH(14, SH_C3[5] * z * (xx - yy));
                DSH(15, SH_C3[6] * x * (xx - 3.f * yy));
                for (int c = 0; c < 3; c++) {
                    dx[c] += (SH_C3[0] * SH(9, c) * 3.f * 2.f * xy + SH_C3[1] * SH(10, c) * yz +
                              SH_C3[2] * SH(11, c) * -2.f * xy + SH_C3[3] * SH(12, c) * -3.f * 2.f * xz +
                              SH_C3[4] * SH(13, c) * (-3.f * xx + 4.f * zz - yy) + SH_C3[5] * SH(14, c) * 2.f * xz +
                              SH_C3[6] * SH(15, c) * 3.f * (xx - yy));
                    dy[c] += (SH_C3[0] * SH(9, c) * 3.f * (xx - yy) + SH_C3[1] * SH(10, c) * xz +
                              SH_C3[2] * SH(11, c) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SH(12, c) * -3.f * 2.f * yz +
                              SH_C3[4] * SH(13, c) * -2.f * xy + SH_C3[5] * SH(14, c) * -2.f * yz +
                              SH_C3[6] * SH(15, c) * -3.f * 2.f * xy);
                    dz[c] += (SH_C3[1] * SH(10, c) * xy + SH_C3[2] * SH(11, c) * 4.f * 2.f * yz +
                              SH_C3[3] * SH(12, c) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SH(13, c) * 4.f * 2.f * xz +
                              SH_C3[5] * SH(14, c) * (xx - yy));
                }
            }
        }
    }
#undef SH
#undef DSH
    float ddx = dx[0] * dRGB[0] + dx[1] * dRGB[1] + dx[2] * dRGB[2];
    float ddy = dy[0] * dRGB[0] + dy[1] * dRGB[1] + dy[2] * dRGB[2];
    float ddz = dz[0] * dRGB[0] + dz[1] * dRGB[1] + dz[2] * dRGB[2];
    /* dnormvdv (auxiliary.h:129-139) */
    float sum2 = dox * dox + doy * doy + doz * doz;
    float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dL_dmeans[3 * (size_t)idx + 0] += ((+sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * invsum32;
    dL_dmeans[3 * (size_t)idx + 1] += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * invsum32;
    dL_dmeans[3 * (size_t)idx + 2] += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * invsum32;
}

void orc_preprocess_backward(int P, int D, int M, const float* means3D, const int* radii,
                             const float* shs, const uint8_t* clamped, const float* scales,
                             const float* rotations, float scale_modifier, const float* transMats,
                             const float* view, const float* proj, int W_true, int H_true,
                             float tan_fovx, float tan_fovy, const float* campos, int flags,
                             float* dL_dmean2D /*[P][3] in/out*/, const float* dL_dnormal3D,
                             float* dL_dtransMat /*in/out*/, const float* dL_dcolors, float* dL_dsh,
                             float* dL_dmean3D, float* dL_dscales, float* dL_drots) {
    (void)scale_modifier; /* Q6: backward ignores scale_modifier (backward.cu:507) */
    int W = W_true, H = H_true;
    if (flags & ORC_FLAG_BWD_WH_QUIRK) {
        const float focal_y = (float)H_true / (2.0f * tan_fovy);
        const float focal_x = (float)W_true / (2.0f * tan_fovx);
        W = f2i_rz((focal_x * tan_fovx) * 2.0f);
        H = f2i_rz((focal_y * tan_fovy) * 2.0f);
    }
    const int precomp = (scales == NULL);
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        if (!(radii[idx] > 0)) continue;
        float T[3][3]; /* T[i] = glm column i = (Tu | Tv | Tw) */
        float Pm[3][4]; /* glm mat3x4 P = world2ndc * ndc2pix: Pm[c][r] */
        float R[3][3], normal[3] = {0, 0, 0};
        float sx = 0, sy = 0;
        const float* p = means3D + 3 * (size_t)idx;
        const float* q = rotations ? rotations + 4 * (size_t)idx : NULL;
        if (precomp) {
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[i][j] = transMats[9 * (size_t)idx + 3 * i + j];
        } else {
            quat_to_rot(q, R[0], R[1], R[2]);
            sx = scales[2 * (size_t)idx]; sy = scales[2 * (size_t)idx + 1];
            float L0[3] = {R[0][0] * sx, R[0][1] * sx, R[0][2] * sx};
            float L1[3] = {R[1][0] * sy, R[1][1] * sy, R[1][2] * sy};
            float Mm[3][4] = {{L0[0], L0[1], L0[2], 0.0f}, {L1[0], L1[1], L1[2], 0.0f}, {p[0], p[1], p[2], 1.0f}};
            const float n2p[3][4] = {{(float)W * 0.5f, 0, 0, (float)(W - 1) * 0.5f},
                                     {0, (float)H * 0.5f, 0, (float)(H - 1) * 0.5f},
                                     {0, 0, 0, 1.0f}};
            /* world2ndc as math matrix A[r][c] = proj[4r + c];  P = A * n2p  (4x3) */
            for (int c = 0; c < 3; c++)
                for (int r = 0; r < 4; r++) {
                    float s = 0;
                    for (int k = 0; k < 4; k++) s += proj[4 * r + k] * n2p[c][k];
                    Pm[c][r] = s;
                }
            /* T = transpose(M) * P : T[c][i] = sum_k M[i][k] * P[c][k] */
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
        const float dm2x = dL_dmean2D[3 * (size_t)idx], dm2y = dL_dmean2D[3 * (size_t)idx + 1];
        int early_return = 0;
        if (dm2x != 0 || dm2y != 0) {
            const float tv[3] = {9.0f, 9.0f, -1.0f};
            float d = tv[0] * T[2][0] * T[2][0] + tv[1] * T[2][1] * T[2][1] + tv[2] * T[2][2] * T[2][2];
            float fv[3], dT3[3], df[3];
            for (int j = 0; j < 3; j++) fv[j] = tv[j] * (1.0f / d);
            for (int j = 0; j < 3; j++) {
                dT[0][j] += dm2x * fv[j] * T[2][j];
                dT[1][j] += dm2y * fv[j] * T[2][j];
                dT3[j] = dm2x * fv[j] * T[0][j] + dm2y * fv[j] * T[1][j];
                df[j] = dm2x * T[0][j] * T[2][j] + dm2y * T[1][j] * T[2][j];
            }
            float dL_dd = (df[0] * fv[0] + df[1] * fv[1] + df[2] * fv[2]) * (-1.0f / d);
            for (int j = 0; j < 3; j++) dT[2][j] += dT3[j] + dL_dd * (tv[j] * T[2][j] * 2.0f);
            if (precomp) {
                for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dL_dtransMat[9 * (size_t)idx + 3 * i + j] = dT[i][j];
                early_return = 1;
            }
        }
        if (!precomp && !early_return) {
            /* dL_dM = P * transpose(dL_dT): dM[c][r] = sum_k P[k][r] * dT[k][c] */
            float dM[3][4];
            for (int c = 0; c < 3; c++)
                for (int r = 0; r < 4; r++) dM[c][r] = Pm[0][r] * dT[0][c] + Pm[1][r] * dT[1][c] + Pm[2][r] * dT[2][c];
            const float* dn = dL_dnormal3D + 3 * (size_t)idx;
            /* transformVec4x3Transpose (auxiliary.h:111-119) */
            float dtn[3] = {view[0] * dn[0] + view[1] * dn[1] + view[2] * dn[2],
                            view[4] * dn[0] + view[5] * dn[1] + view[6] * dn[2],
                            view[8] * dn[0] + view[9] * dn[1] + view[10] * dn[2]};
            float pvx = view[0] * p[0] + view[4] * p[1] + view[8] * p[2] + view[12];
            float pvy = view[1] * p[0] + view[5] * p[1] + view[9] * p[2] + view[13];
            float pvz = view[2] * p[0] + view[6] * p[1] + view[10] * p[2] + view[14];
            float cosv = -(pvx * normal[0] + pvy * normal[1] + pvz * normal[2]);
            float mult = cosv > 0 ? 1.0f : -1.0f;
            for (int j = 0; j < 3; j++) dtn[j] *= mult;
            float dRS[3][3] = {{dM[0][0], dM[0][1], dM[0][2]}, {dM[1][0], dM[1][1], dM[1][2]}, {dtn[0], dtn[1], dtn[2]}};
            float vR[3][3];
            for (int j = 0; j < 3; j++) { vR[0][j] = dRS[0][j] * sx; vR[1][j] = dRS[1][j] * sy; vR[2][j] = dRS[2][j]; }
            /* quat_to_rotmat_vjp (auxiliary.h:239-283) -- uses rsqrtf in the reference */
            float s = 1.0f / sqrtf(q[3] * q[3] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
            float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
            float* dq = dL_drots + 4 * (size_t)idx;
            dq[0] = 2.f * (x * (vR[1][2] - vR[2][1]) + y * (vR[2][0] - vR[0][2]) + z * (vR[0][1] - vR[1][0]));
            dq[1] = 2.f * (-2.f * x * (vR[1][1] + vR[2][2]) + y * (vR[0][1] + vR[1][0]) + z * (vR[0][2] + vR[2][0]) + w * (vR[1][2] - vR[2][1]));
            dq[2] = 2.f * (x * (vR[0][1] + vR[1][0]) - 2.f * y * (vR[0][0] + vR[2][2]) + z * (vR[1][2] + vR[2][1]) + w * (vR[2][0] - vR[0][2]));
            dq[3] = 2.f * (x * (vR[0][2] + vR[2][0]) + y * (vR[1][2] + vR[2][1]) - 2.f * z * (vR[0][0] + vR[1][1]) + w * (vR[0][1] - vR[1][0]));
            dL_dscales[2 * (size_t)idx + 0] = dRS[0][0] * R[0][0] + dRS[0][1] * R[0][1] + dRS[0][2] * R[0][2];
            dL_dscales[2 * (size_t)idx + 1] = dRS[1][0] * R[1][0] + dRS[1][1] * R[1][1] + dRS[1][2] * R[1][2];
            dL_dmean3D[3 * (size_t)idx + 0] = dM[2][0];
            dL_dmean3D[3 * (size_t)idx + 1] = dM[2][1];
            dL_dmean3D[3 * (size_t)idx + 2] = dM[2][2];
        }
        if (shs) sh_backward(idx, D, M, means3D, campos, shs, clamped, dL_dcolors, dL_dmean3D, dL_dsh);
        /* densification proxy (backward.cu:652-655); reads dL_dtransMats AFTER the precomp write-back */
        float depth = transMats[9 * (size_t)idx + 8];
        dL_dmean2D[3 * (size_t)idx + 0] = dL_dtransMat[9 * (size_t)idx + 2] * depth * 0.5f * (float)W;
        dL_dmean2D[3 * (size_t)idx + 1] = dL_dtransMat[9 * (size_t)idx + 5] * depth * 0.5f * (float)H;
    }
}

/* ------------------------------------------------------------------------------------------
 * distCUDA2 (submodules/simple-knn/simple_knn.cu:120-184): mean of the 3 smallest squared
 * distances to other points.  Brute force (the Morton/box structure is only an accelerator).
 * ---------------------------------------------------------------------------------------- */
void orc_knn_mean_dist2(int P, const float* pts, float* out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        float best[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
        const float rx = pts[3 * (size_t)i], ry = pts[3 * (size_t)i + 1], rz = pts[3 * (size_t)i + 2];
        for (int j = 0; j < P; j++) {
            if (j == i) continue;
            float dx = pts[3 * (size_t)j] - rx, dy = pts[3 * (size_t)j + 1] - ry, dz = pts[3 * (size_t)j + 2] - rz;
            float dist = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            for (int k = 0; k < 3; k++)
                if (best[k] > dist) { float t = best[k]; best[k] = dist; dist = t; }
        }
        out[i] = ((best[0] + best[1]) + best[2]) / 3.0f;
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
