// Stage "c2f": per finer level  x2 bilinear upsample (x2.0)  ->  3x3-candidate x 4-affine-model plane-fitting refine
// ->  21x21 joint-bilateral flow smoothing; one more smoothing at level 0.
// Restates baoCudaBLF_C2F (bao_pmflow_refine_kernel.cu:1076-1087), _d_bao_bilinear_resize<float2> and
// _d_bao_multiply_scalar (basic/bao_basic_cuda.cuh:511-537,135-149), d_bilateral_refine_flow_planefitting and
// _d_compute_patch_dist_planefitting (bao_pmflow_kernel.cu:2005-2041,319-513), d_flow_bilateral_filtering
// (bao_pmflow_refine_kernel.cu:752-826) and the level loop of compute_flow (…cuda.cpp:275-289).
// The weighted-median call inside that loop (…cuda.cpp:281) filters a buffer nothing reads and is not executed.
//
// B200 design
//  * upsample, x2 scale and refine are ONE kernel: a refine thread only needs the upsampled flow of its own pixel.
//  * refine: one thread per (pixel, candidate column m); its 3 candidates x 4 models accumulate side by side while the
//    image-1 side of every sample (colour, census, range distance d1, spatial weight) is computed once and reused by
//    all 12 (candidate, model) pairs.  Each pair still adds its 100 samples in the reference's order, so its cost is
//    the bit pattern the reference gets.  The three column results of a pixel meet in shared memory and keep
//    the reference's m-outer / n-inner first-minimum order.  (One thread per pixel with all 36 pairs was measured 33 % slower:
//    168 registers, a third of the warps.)
//  * smoothing reads a snapshot and writes a second buffer (the reference filters in place).
#include <cuda.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include "eppm_internal.h"

namespace eppm {

// plane-fitting coefficient sets (bao_pmflow_kernel.cu:319-332); model 0 is the identity
__constant__ float c_pf[3][4] = {
    {0.177f, -0.011f, -0.003f, 0.301f},   // COEF_FL_{U_X,U_Y,V_X,V_Y}
    {0.125f, -0.357f, 0.009f, 0.308f},    // COEF_LEFT_*
    {0.205f, 0.370f, 0.011f, 0.296f},     // COEF_RIGHT_*
};

// _d_bao_bilinear_resize<float2> at ratio 2 followed by x2.0 (basic/bao_basic_cuda.cuh:511-537, 135-149)
__device__ __forceinline__ float2 upsample2(const float2* __restrict__ src, int ws, int hs, int x, int y) {
    const float div_scale = 0.5f;  // 1.f/ratio with ratio = PYR_RATIO_UP = 2
    const float fx = __fmaf_rn((float)(x + 1), div_scale, -1.f), fy = __fmaf_rn((float)(y + 1), div_scale, -1.f);
    const int xx = (int)fx, yy = (int)fy;
    const float dx = fmaxf(fminf(__fsub_rn(fx, (float)xx), 1.f), 0.f), dy = fmaxf(fminf(__fsub_rn(fy, (float)yy), 1.f), 0.f);
    float rx = 0.f, ry = 0.f;
#pragma unroll
    for (int m = 0; m <= 1; m++)
#pragma unroll
        for (int n = 0; n <= 1; n++) {
            const int u = max(0, min(ws - 1, xx + m)), v = max(0, min(hs - 1, yy + n));
            const float s = __fmul_rn(fabsf(__fsub_rn((float)(1 - m), dx)), fabsf(__fsub_rn((float)(1 - n), dy)));
            const float2 p = src[(size_t)v * ws + u];
            rx = __fmaf_rn(s, p.x, rx);
            ry = __fmaf_rn(s, p.y, ry);
        }
    return make_float2(__fmul_rn(rx, 2.0f), __fmul_rn(ry, 2.0f));
}

struct RefineArgs {
    const float4* pix1;   // PADDED origin of pair 0, image 1 / image 2 at this level
    const float4* pix2;
    size_t plane;
    int pw, w, h;
    const float2* coarse; // [B][hs][ws]
    int ws, hs;
    float2* flow;         // [B][h][w] out
    int upsample;         // 1: coarse is the next-coarser level (x2 upsample fused); 0: coarse is already at this level
    int y0;               // first row of the band this launch owns (grid.y = rows in the band)
    int px_bytes;         // sizeof(float4) as a run-time value: the product off * px_bytes is then formed once per site and shared by the three
                          // candidate rows (a literal 16 is strength-reduced per load into a shift, a high-part multiply and a 64-bit add)
};

// CTA = 3 warps x 32 pixels: warp m evaluates candidate column m of 32 consecutive pixels of one row (coalesced plane
// reads); the three column minima of a pixel meet in shared memory.
constexpr int RF_PIX = 32;
__device__ __forceinline__ float min_ref(float a, float b) { return a < b ? a : b; }  // the reference's __min macro
#ifndef RF_MINBLOCKS
#define RF_MINBLOCKS 6  // 96 registers -> 6 CTAs (18 warps) per SM; measured faster than 128 registers / 5 CTAs
#endif
template <int STRIDE>
__global__ void __launch_bounds__(RF_PIX * 3, RF_MINBLOCKS) k_c2f_refine(RefineArgs a, const __grid_constant__ CostLut lut) {
    __shared__ float s_best[3][RF_PIX];
    __shared__ int s_bn[3][RF_PIX];
    __shared__ float s_census[CENSUS_LUT_N];
    load_census_lut(s_census, lut);
    const int m = threadIdx.x >> 5, pl = threadIdx.x & 31;
    const int x = blockIdx.x * RF_PIX + pl, y = a.y0 + blockIdx.y;
    const bool in = x < a.w;
    const int b = blockIdx.z;
    const float4* I1 = a.pix1 + (size_t)b * a.plane;
    const float4* I2 = a.pix2 + (size_t)b * a.plane;
    asm volatile("" : "+l"(I1), "+l"(I2));  // keep the per-pair plane bases in registers: every load is then base + u32 offset
    float2 fl = make_float2(0.f, 0.f);
    if (in) fl = a.upsample ? upsample2(a.coarse + (size_t)b * a.ws * a.hs, a.ws, a.hs, x, y) : a.coarse[(size_t)b * a.w * a.h + (size_t)y * a.w + x];
    // :2011 unknown flow -> 0 and done
    const bool unknown = fl.x > EPPM_UNKNOWN_FLOW_THRESH || fl.y > EPPM_UNKNOWN_FLOW_THRESH;
    // :2014-2019 candidates: short(flow)+pos, -1/+0/+1 (16-bit arithmetic as in the reference)
    const short cxc = (short)((short)(int)fl.x + x), cyc = (short)((short)(int)fl.y + y);
    const short cx = (short)(cxc + (m - 1));
    float cost[3] = {FLT_MAX, FLT_MAX, FLT_MAX};  // best-of-4-models cost of candidate rows n = 0..2
    bool valid[3];
#pragma unroll
    for (int n = 0; n < 3; n++) {
        const short cy = (short)(cyc + (n - 1));
        valid[n] = in && !unknown && !(cx < 0 || cy < 0 || cx >= a.w || cy >= a.h);  // :2029
    }
    if (valid[0] || valid[1] || valid[2]) {
        // accumulators: [n][model]
        float cs[3][4], ws[3][4];
#pragma unroll
        for (int n = 0; n < 3; n++)
#pragma unroll
            for (int q = 0; q < 4; q++) cs[n][q] = ws[n][q] = 0.f;
        const float4* a0 = I1 + (unsigned)((y + PAD) * a.pw + x + PAD);
        const PixPk c1k = pack_pix(ldpix(a0));
        PixPk c2k[3];
#pragma unroll
        for (int n = 0; n < 3; n++) {
            const int cy = max(-PAD, min(a.h - 1 + PAD, (int)cyc + n - 1));  // invalid rows are never used; clamp keeps the load in-plane
            const int cxs = max(-PAD, min(a.w - 1 + PAD, (int)cx));
            c2k[n] = pack_pix(ldpix(I2 + (unsigned)((cy + PAD) * a.pw + cxs + PAD)));
        }
        const float uu = (float)((int)cx - x);  // :350 float uu = x2 - x1
#pragma unroll 1
        for (int i = -PATCH_R; i <= PATCH_R; i += STRIDE) {
            const float fi = (float)i;
            const int ai = i < 0 ? -i : i;
#pragma unroll 2
            for (int j = -PATCH_R; j <= PATCH_R; j += STRIDE) {
                const float fj = (float)j;
                const float4 p1 = ldpix(a0 + i * a.pw + j);
                const PixPk p1k = pack_pix(p1);
                const float d1 = max3abs_diff(c1k, p1k);
                const float gg = lut.gg[ai][j < 0 ? -j : j];
                // x coordinates of the 4 models: cx2 = fma(i, C_uy, fma(j, C_ux, float(x1+j) + uu))   (:402, :440, :478 as contracted)
                const float bx = __fadd_rn(uu, (float)(x + j));
                int sx[4];
                sx[0] = (int)cx + j + PAD;  // identity model: float(x1+j) + float(x2-x1) is the exact integer x2+j
#pragma unroll
                for (int q = 0; q < 3; q++) sx[q + 1] = __float2int_rd(__fmaf_rn(fi, c_pf[q][1], __fmaf_rn(fj, c_pf[q][0], bx))) + PAD;
#pragma unroll
                for (int n = 0; n < 3; n++) {
                    if (!valid[n]) continue;
                    const int cy = (int)cyc + n - 1;
                    const float vv = (float)(cy - y);
                    const float by = __fadd_rn((float)(y + i), vv);
                    int sy[4];
                    sy[0] = cy + i + PAD;
#pragma unroll
                    for (int q = 0; q < 3; q++) sy[q + 1] = __float2int_rd(__fmaf_rn(fi, c_pf[q][3], __fmaf_rn(fj, c_pf[q][2], by))) + PAD;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const float4 p2 = ldpix(I2 + (unsigned)(sy[q] * a.pw + sx[q]));
                        sample_term(p1, p1k, p2, c2k[n], d1, gg, s_census, cs[n][q], ws[n][q]);
                    }
                }
            }
        }
#pragma unroll
        for (int n = 0; n < 3; n++) {
            if (!valid[n]) continue;
            const float k1 = __fdiv_rn(cs[n][0], ws[n][0]), k2 = __fdiv_rn(cs[n][1], ws[n][1]);
            const float k3 = __fdiv_rn(cs[n][2], ws[n][2]), k4 = __fdiv_rn(cs[n][3], ws[n][3]);
            cost[n] = min_ref(k1, min_ref(k2, min_ref(k3, k4)));  // :512 __min(cost1,__min(cost2,__min(cost3,cost4)))
        }
    }
    // arg-min in the reference's order: m outer, n inner, strict '<' against 999999 (:2024,:2031)
    float best = 999999.f;
    int best_n = -1;
#pragma unroll
    for (int n = 0; n < 3; n++)
        if (valid[n] && cost[n] < best) { best = cost[n]; best_n = n; }
    s_best[m][pl] = best;
    s_bn[m][pl] = best_n;
    __syncthreads();
    float bcost = 999999.f;
    int bm = -1, bn = -1;
#pragma unroll
    for (int mm = 0; mm < 3; mm++) {
        const float oc = s_best[mm][pl];
        const int on = s_bn[mm][pl];
        if (on >= 0 && oc < bcost) { bcost = oc; bm = mm; bn = on; }
    }
    if (in && m == 0) {
        float2 out;
        if (unknown) out = make_float2(0.f, 0.f);
        else {
            short bx = cxc, by = cyc;  // :2020-2022 default = centre candidate
            if (bm >= 0) { bx = (short)(cxc + (bm - 1)); by = (short)(cyc + (bn - 1)); }
            out = make_float2((float)(bx - x), (float)(by - y));  // :2038-2039
        }
        a.flow[(size_t)b * a.w * a.h + (size_t)y * a.w + x] = out;
    }
}

// ---- table-driven refine (default at patch stride 2) ----
// The sample sites of the three affine models are floor(fma(i, C_y, fma(j, C_x, float(X)))) with X = cx + j an exact integer
// (bao_pmflow_kernel.cu:402,440,478 as contracted by nvcc).  For odd (i, j) the real value j*C_x + i*C_y never comes closer than
// 1e-3 to an integer, far more than the two FFMA roundings can move it at any |X| < 2^15, so site - X is a function of (i, j, model)
// alone.  The same holds at stride 3; at stride 1 one site (model 3, (i, j) = (-7, -2): -2*0.205 - 7*0.370 = -3.00000003) does depend
// on X: the table holds that site without its dx and the kernel computes dx with the reference's two FFMAs (AffineTab::exc_*).  build_affine_tab() tabulates it and CHECKS that claim for every X the level can produce with the host's correctly rounded
// fmaf; only then is this kernel used.  It replaces 2 FFMA + F2I + IADD + IMAD per coordinate by one table read per site,
// and groups the `t2 < -126` fix-up of __expf (taken by ~1 % of the samples) of the four models of a candidate into one test.
// base + off pixels as ONE IMAD.WIDE (left to itself the compiler sign-extends and shifts with three ALU instructions)
__device__ __forceinline__ const float4* pix_at(const float4* base, int off) {
    const float4* r;
    asm("mad.wide.s32 %0, %1, 16, %2;" : "=l"(r) : "r"(off), "l"(base));
    return r;
}

// base + off pixels with the multiplier (16) taken from kernel parameter space, where ptxas cannot fold it (see RefineArgs::px_bytes)
__device__ __forceinline__ const float4* pix_at_b(const float4* base, int off, int px_bytes) {
    const float4* r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(off), "r"(px_bytes), "l"(base));
    return r;
}

// NCT = candidate rows per thread.  3: CTA = 3 warps, warp m owns candidate column m (96 registers, 18 warps per SM).
// 1: CTA = 9 warps, warp (m, n) owns ONE candidate and its four models (72 registers, 27 warps per SM): the image-1 side of a
// sample is shared by 4 instead of 12 accumulator pairs (33.4 instead of 30.7 instructions per sample).  Measured equal (8.70 vs 8.74 ms
// per 1080p pair at level 0; 56 registers / 36 warps: 9.27 ms) -- the kernel is not occupancy bound; kept behind EPPM_VARIANT=256.
#ifndef RF_TAB2_MINBLOCKS
#define RF_TAB2_MINBLOCKS 7   // the default (stride-2, table) instantiation: 80 registers, 7 CTAs = 21 warps per SM, 80 B of spill outside the sample
                              // loop; measured 8.46 ms per 1080p pair at level 0 against 8.75 (6 CTAs, 96 registers), 9.26 (5), 8.83 (8)
#endif
#ifndef RF_JUNROLL
#define RF_JUNROLL 2   // samples of a patch row handled per iteration of the inner loop (tuning knob)
#endif
#ifndef S4_UNROLL
#define S4_UNROLL 7     // taps of a tile row per iteration of k_flow_smooth4 (tuning knob; final smoothing per pair: 1: 0.577, 3: 0.515, 7: 0.507 ms)
#endif
#ifndef RF_ROW_JUNROLL
#define RF_ROW_JUNROLL 1   // the same for k_c2f_refine_row (default kernel)
#endif
#define EPPM_PRAGMA_(x) _Pragma(#x)
#define EPPM_PRAGMA(x) EPPM_PRAGMA_(x)
template <bool GROUP_TINY, int NCT, int MINB, int STRIDE, bool ALLROWS = false, bool WIDE = false>
__global__ void __launch_bounds__(RF_PIX * 9 / NCT, MINB)
    k_c2f_refine_tab(RefineArgs a, const __grid_constant__ CostLut lut, const __grid_constant__ AffineTab tab) {
    __shared__ float s_best[9][RF_PIX];
    __shared__ float s_census[CENSUS_LUT_N];
    load_census_lut(s_census, lut);
    const unsigned lut_base = census_lut_base(s_census);
    const int wq = threadIdx.x >> 5, pl = threadIdx.x & 31;
    const int m = NCT == 3 ? wq : wq / 3, n0 = NCT == 3 ? 0 : wq - 3 * m;
    const int x = blockIdx.x * RF_PIX + pl, y = a.y0 + blockIdx.y;
    const bool in = x < a.w;
    const int b = blockIdx.z;
    const float4* I1 = a.pix1 + (size_t)b * a.plane;
    const float4* I2 = a.pix2 + (size_t)b * a.plane;
    float2 fl = make_float2(0.f, 0.f);
    if (in) fl = a.upsample ? upsample2(a.coarse + (size_t)b * a.ws * a.hs, a.ws, a.hs, x, y) : a.coarse[(size_t)b * a.w * a.h + (size_t)y * a.w + x];
    const bool unknown = fl.x > EPPM_UNKNOWN_FLOW_THRESH || fl.y > EPPM_UNKNOWN_FLOW_THRESH;  // :2011
    const short cxc = (short)((short)(int)fl.x + x), cyc = (short)((short)(int)fl.y + y);     // :2014-2019
    const short cx = (short)(cxc + (m - 1));
    float cost[NCT];
    bool valid[NCT];
    bool any = false;
#pragma unroll
    for (int n = 0; n < NCT; n++) {
        const short cy = (short)(cyc + (n0 + n - 1));
        cost[n] = FLT_MAX;
        valid[n] = in && !unknown && !(cx < 0 || cy < 0 || cx >= a.w || cy >= a.h);  // :2029
        any = any || valid[n];
    }
    if (any) {
        float cs[NCT][4], ws[NCT][4];
#pragma unroll
        for (int n = 0; n < NCT; n++)
#pragma unroll
            for (int q = 0; q < 4; q++) cs[n][q] = ws[n][q] = 0.f;
        const float4* a0 = I1 + (unsigned)((y + PAD) * a.pw + x + PAD);
        const PixPk c1k = pack_pix(ldpix(a0));
        PixPk c2k[NCT];
        const float4* P[NCT];   // candidate centres in image 2 (rows that are not valid are clamped into the plane and never used)
#pragma unroll
        for (int n = 0; n < NCT; n++) {
            const int cy = max(0, min(a.h - 1, (int)cyc + n0 + n - 1));   // rows that are not valid are scored at a clamped centre and never used
            const int cxs = max(0, min(a.w - 1, (int)cx));
            P[n] = I2 + (unsigned)((cy + PAD) * a.pw + cxs + PAD);
            c2k[n] = pack_pix(ldpix(P[n]));
            asm volatile("" : "+l"(P[n]));  // keep the centre pointers in registers: every site is then one IMAD.WIDE away
        }
        int s = 0;
#pragma unroll 1
        for (int i = -PATCH_R; i <= PATCH_R; i += STRIDE) {
            const int ai = i < 0 ? -i : i;
            const int irow = i * a.pw;
EPPM_PRAGMA(unroll RF_JUNROLL)
            for (int j = -PATCH_R; j <= PATCH_R; j += STRIDE, s++) {
                const float4 p1 = ldpix(a0 + irow + j);
                const PixPk p1k = pack_pix(p1);
                const float d1 = max3abs_diff(c1k, p1k);
                const float gg = lut.gg[ai][j < 0 ? -j : j];
                int off[4];
                off[0] = irow + j;  // identity model: the exact integer site (cx + j, cy + i)
#pragma unroll
                for (int q = 0; q < 3; q++) off[q + 1] = tab.off[q][s];
                if (STRIDE == 1 && s == tab.exc_s) {   // the one site of stride 1 whose x offset depends on the coordinate itself (see AffineTab)
                    const int X = (int)cx + j;
                    const int dx = __float2int_rd(__fmaf_rn((float)i, tab.exc_ci, __fmaf_rn((float)j, tab.exc_cj, (float)X))) - X;
#pragma unroll
                    for (int q = 0; q < 3; q++)
                        if (q == tab.exc_q) off[q + 1] += dx;
                }
#pragma unroll
                for (int n = 0; n < NCT; n++) {
#ifndef RF_NOVALID
                    // ALLROWS: candidate rows outside the image (:2029) are scored at a clamped centre and discarded instead of being branched
                    // around: only warps at the image border contain such rows, the divergent branch costs BSSY + BSYNC + BRA per candidate row
                    // and sample everywhere else
                    if (NCT > 1 && !ALLROWS && !valid[n]) continue;
#endif
                    if (GROUP_TINY) {
                        float ct[4], t2[4], w[4];
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            sample_eval(p1, p1k, ldpix(WIDE ? pix_at_b(P[n], off[q], a.px_bytes) : pix_at(P[n], off[q])), c2k[n], d1, lut_base, ct[q], t2[q]);
#pragma unroll
                        for (int q = 0; q < 4; q++) w[q] = __fmul_rn(ex2_mufu(t2[q]), gg);
                        if (fminf(fminf(t2[0], t2[1]), fminf(t2[2], t2[3])) < -126.0f) {
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                if (t2[q] < -126.0f) w[q] = __fmul_rn(ex2_tiny(t2[q]), gg);
                        }
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            cs[n][q] = __fmaf_rn(ct[q], w[q], cs[n][q]);
                            ws[n][q] = __fadd_rn(ws[n][q], w[q]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; q++) sample_term(p1, p1k, ldpix(pix_at(P[n], off[q])), c2k[n], d1, gg, s_census, cs[n][q], ws[n][q]);
                    }
                }
            }
        }
#pragma unroll
        for (int n = 0; n < NCT; n++) {
            if (!valid[n]) continue;
            const float k1 = __fdiv_rn(cs[n][0], ws[n][0]), k2 = __fdiv_rn(cs[n][1], ws[n][1]);
            const float k3 = __fdiv_rn(cs[n][2], ws[n][2]), k4 = __fdiv_rn(cs[n][3], ws[n][3]);
            cost[n] = min_ref(k1, min_ref(k2, min_ref(k3, k4)));  // :512
        }
    }
    // candidates that are not valid never win: the reference skips them (:2029), here they carry +inf against the strict '<' below
#pragma unroll
    for (int n = 0; n < NCT; n++) s_best[m * 3 + n0 + n][pl] = valid[n] ? cost[n] : __int_as_float(0x7f800000);
    __syncthreads();
    if (in && wq == 0) {
        float2 out;
        if (unknown) out = make_float2(0.f, 0.f);
        else {
            // arg-min in the reference's order: m outer, n inner, strict '<' against 999999 (:2024,:2031)
            float bcost = 999999.f;
            int bk = -1;
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const float oc = s_best[k][pl];
                if (oc < bcost) { bcost = oc; bk = k; }
            }
            short bx = cxc, by = cyc;   // :2020-2022 default = centre candidate
            if (bk >= 0) { bx = (short)(cxc + (bk / 3 - 1)); by = (short)(cyc + (bk % 3 - 1)); }
            out = make_float2((float)(bx - x), (float)(by - y));  // :2038-2039
        }
        a.flow[(size_t)b * a.w * a.h + (size_t)y * a.w + x] = out;
    }
}

// ---- table-driven refine, warp = candidate ROW ----
// k_c2f_refine_tab gives warp m the candidate COLUMN m and a thread its three candidate rows: the twelve image-2 sites of a sample are
// twelve 64-bit address computations (two ALU instructions each: 2.3 of the 30.9 issued instructions per sample).  Here warp n owns the
// candidate ROW n and a thread its three candidate columns: the three columns of a (row, model) are neighbouring pixels, base -16 / +0 /
// +16 bytes, which the load instruction takes as an immediate -- four address computations per sample instead of twelve.  Everything
// else (site table, sample order, grouped fix-up, strict '<' arg-min in the reference's order) is k_c2f_refine_tab's.
constexpr int RF_TILE_W = RF_PIX + 2 * PATCH_R, RF_TILE_H = PATCH_R + 1;   // image-1 tile of a CTA at stride 2: 50 pixels x 10 sampled rows (8000 bytes)
// the sample loop of k_c2f_refine_row.  CHECK = false: every candidate of every lane of the warp is valid (all warps but those at the image
// border), so the per-candidate divergence guard (BSSY / BSYNC / BRA: 1.5 of ~30 issued instructions per sample) is not compiled in.
// TINY = false: the `t2 < -126` fix-up of __expf is not compiled in (a bare MUFU.EX2.ftz returns 0 there): only valid for rows behind a prefix
// after which every accumulator is large enough to absorb any such weight unchanged (see k_c2f_refine_row, FASTW).  Rows [i_lo, i_hi].
template <int STRIDE, bool CHECK, class LutRef, bool TILE = false, bool TINY = true>
__device__ __forceinline__ void refine_row_loop(const RefineArgs& a, const CostLut& lut, const AffineTab& tab, const float4* a0, const float4* Pc, const PixPk& c1k,
                                                const PixPk (&c2k)[3], const bool (&valid)[3], unsigned wmask, LutRef lut_ref, float (&cs)[3][4], float (&ws)[3][4],
                                                const float4* tile = nullptr, int i_lo = -PATCH_R, int i_hi = PATCH_R) {
    int s = ((i_lo + PATCH_R) / STRIDE) * ((2 * PATCH_R) / STRIDE + 1);
#pragma unroll 1
    for (int i = i_lo; i <= i_hi; i += STRIDE) {
        const int ai = i < 0 ? -i : i;
        const int irow = i * a.pw;
        // one sample per iteration: the kernel is sensitive to its instruction footprint (measured at level 0: 1: 7.12, 2: 7.16, 5: 8.92, 10: 10.95 ms per pair);
        // hoisting the tile row pointer by hand (7.20) or dropping the scheduling fence below in the guard-free loop (7.43) did not help
EPPM_PRAGMA(unroll RF_ROW_JUNROLL)
        for (int j = -PATCH_R; j <= PATCH_R; j += STRIDE, s++) {
            // TILE: the image-1 samples of the CTA's 32 pixels (every second row of a (32 + 18)-pixel strip) were staged in shared memory by TMA
            const float4 p1 = TILE ? tile[((i + PATCH_R) / STRIDE) * RF_TILE_W + (j + PATCH_R)] : ldpix(a0 + irow + j);
            const PixPk p1k = pack_pix(p1);
            const float d1 = max3abs_diff(c1k, p1k);
            const float gg = (STRIDE == 2 && TILE) ? tab.gs[s] : lut.gg[ai][j < 0 ? -j : j];
            int off[4];
            off[0] = irow + j;  // identity model: the exact integer site (cx + j, cy + i)
#pragma unroll
            for (int q = 0; q < 3; q++) off[q + 1] = tab.off[q][s];
            const float4* site[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
                site[q] = (STRIDE == 2 && TILE) ? reinterpret_cast<const float4*>(reinterpret_cast<const char*>(Pc) + tab.boff[q][s]) : pix_at(Pc, off[q]);
#pragma unroll
            for (int m = 0; m < 3; m++) {
                if (CHECK && !valid[m]) continue;
                // CHECK = false: wmask is warp-uniform and all ones; the (uniform, never divergent) test keeps the three candidates in separate
                // basic blocks -- scheduled as one block they need more registers than the kernel has
#ifndef RF_FENCE_FROM
#define RF_FENCE_FROM 0
#endif
                if (!CHECK && m >= RF_FENCE_FROM && !(wmask & (1u << m))) continue;   // (a __syncwarp() as the fence instead: 7.50 vs 7.17 ms per pair at level 0)
                float ct[4], t2[4], w[4];
#pragma unroll
                for (int q = 0; q < 4; q++) sample_eval(p1, p1k, ldpix(site[q] + (m - 1)), c2k[m], d1, lut_ref, ct[q], t2[q]);
#pragma unroll
                for (int q = 0; q < 4; q++) w[q] = __fmul_rn(ex2_mufu(t2[q]), gg);
                if (TINY && fminf(fminf(t2[0], t2[1]), fminf(t2[2], t2[3])) < -126.0f) {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (t2[q] < -126.0f) w[q] = __fmul_rn(ex2_tiny(t2[q]), gg);
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    cs[m][q] = __fmaf_rn(ct[q], w[q], cs[m][q]);
                    ws[m][q] = __fadd_rn(ws[m][q], w[q]);
                }
            }
        }
    }
}

// LUT0: the census table lives at the user base of the shared window (see Lut0); FAST: warps whose 96 candidates are all valid take a loop
// without validity guards.
// FASTW (with TMA1): the first patch row is scored with the exact __expf fix-up; if afterwards EVERY accumulator of every lane's valid candidates is
// at least 2^-99, the other rows run without the fix-up test (1.5 of ~28.7 issued instructions per sample).  Exact, not approximate: a weight that
// needs the fix-up is below 2^-126 (its cost term below 2^-125), the sums never decrease (all terms are non-negative), and adding a
// non-negative value below half an ulp of a sum leaves the sum unchanged under round-to-nearest -- so both forms produce the same bits.
// Warps that fail the test (static scenes: cost sums exactly 0; pixels boxed in by saturated edges) keep the exact loop.
template <int MINB, int STRIDE, bool LUT0, bool FAST, int LUTX = 0, int UNI = 0, bool TMA1 = false, bool FASTW = false>
__global__ void __launch_bounds__(RF_PIX * 3, MINB)
    k_c2f_refine_row(RefineArgs a, const __grid_constant__ CostLut lut, const __grid_constant__ AffineTab tab, const __grid_constant__ CUtensorMap tmap1) {
    // no static shared memory in this kernel.  LUTX = 0: [0, 16) census table by popcount (9 used), then s_best[9][RF_PIX];
    // LUTX = REP: [0, 256 * REP) census table by XOR byte, replicated (see LutX), then s_best
    extern __shared__ float s_dyn[];
    float* s_census = s_dyn;
    constexpr int LUT_WORDS = LUTX ? 256 * LUTX : 16;
    float (*s_best)[RF_PIX] = reinterpret_cast<float (*)[RF_PIX]>(s_dyn + LUT_WORDS);
    if (LUTX) {
        for (int k = threadIdx.x; k < 256 * LUTX; k += blockDim.x) s_dyn[k] = lut.census[__popc((unsigned)k / LUTX)];
        __syncthreads();
    } else {
        load_census_lut(s_census, lut);
    }
    const unsigned lut_base = census_lut_base(s_census);
    const int n = threadIdx.x >> 5, pl = threadIdx.x & 31;   // n: candidate row of this warp
    const int x = blockIdx.x * RF_PIX + pl, y = a.y0 + blockIdx.y;
    const bool in = x < a.w;
    const int b = blockIdx.z;
    // TMA1 (default): the image-1 samples of the CTA -- columns x0-9 .. x0+40, rows y-9, y-7, .. y+9 of the packed plane -- as ONE bulk tensor
    // copy with an element stride of 2 in y (cp.async.bulk.tensor + mbarrier), instead of one L1-resident 16-byte load per sample and thread
    float4* s_tile = reinterpret_cast<float4*>(reinterpret_cast<char*>(s_dyn) + 1280);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(s_dyn) + 1280 + RF_TILE_W * RF_TILE_H * sizeof(float4));
    if (TMA1) {
        if (threadIdx.x == 0) {
            mbar_init(s_bar, 1);
            mbar_expect_tx(s_bar, RF_TILE_W * RF_TILE_H * sizeof(float4));
            tma_load_3d(s_tile, &tmap1, (blockIdx.x * RF_PIX + PAD - PATCH_R) * 4, y + PAD - PATCH_R, b, s_bar);
        }
    }
    const float4* I1 = a.pix1 + (size_t)b * a.plane;
    const float4* I2 = a.pix2 + (size_t)b * a.plane;
    float2 fl = make_float2(0.f, 0.f);
    if (in) fl = a.upsample ? upsample2(a.coarse + (size_t)b * a.ws * a.hs, a.ws, a.hs, x, y) : a.coarse[(size_t)b * a.w * a.h + (size_t)y * a.w + x];
    const bool unknown = fl.x > EPPM_UNKNOWN_FLOW_THRESH || fl.y > EPPM_UNKNOWN_FLOW_THRESH;  // :2011
    const short cxc = (short)((short)(int)fl.x + x), cyc = (short)((short)(int)fl.y + y);     // :2014-2019
    const short cy = (short)(cyc + (n - 1));
    if (TMA1) {
        __syncthreads();          // the barrier object is initialised before anyone polls it
        mbar_wait(s_bar, 0);
    }
    float cost[3];
    bool valid[3];
    bool any = false, all = true;
#pragma unroll
    for (int m = 0; m < 3; m++) {
        const short cx = (short)(cxc + (m - 1));
        cost[m] = FLT_MAX;
        valid[m] = in && !unknown && !(cx < 0 || cy < 0 || cx >= a.w || cy >= a.h);  // :2029
        any = any || valid[m];
        all = all && valid[m];
    }
    const bool warp_all = FAST && __all_sync(0xffffffffu, all);
    unsigned wmask = __ballot_sync(0xffffffffu, all) == 0xffffffffu ? 7u : 0u;
    if (UNI) {   // candidate m is scored by the whole warp when ANY lane needs it (lanes that do not discard their result; their sites are clamped into the plane)
        wmask = (__any_sync(0xffffffffu, valid[0]) ? 1u : 0u) | (__any_sync(0xffffffffu, valid[1]) ? 2u : 0u) | (__any_sync(0xffffffffu, valid[2]) ? 4u : 0u);
    }
    asm volatile("" : "+r"(wmask));   // opaque to the compiler: see refine_row_loop
    const unsigned any_mask = FASTW ? __ballot_sync(0xffffffffu, any) : 0u;   // the lanes that enter the scoring branch (FASTW votes among them)
    if (any) {
        float cs[3][4], ws[3][4];
#pragma unroll
        for (int m = 0; m < 3; m++)
#pragma unroll
            for (int q = 0; q < 4; q++) cs[m][q] = ws[m][q] = 0.f;
        const float4* a0 = I1 + (unsigned)((y + PAD) * a.pw + x + PAD);
        const PixPk c1k = pack_pix(ldpix(a0));
        // centre of the MIDDLE candidate of this row; a valid candidate has cx in [0, w), so the middle one lies in [-1, w]: inside the padded plane
        const int cys = max(0, min(a.h - 1, (int)cy)), cxs = max(-1, min(a.w, (int)cxc));
        const float4* Pc = I2 + (unsigned)((cys + PAD) * a.pw + cxs + PAD);
        asm volatile("" : "+l"(Pc));
        PixPk c2k[3];
#pragma unroll
        for (int m = 0; m < 3; m++) c2k[m] = pack_pix(ldpix(Pc + (m - 1)));
        if (TMA1 && FASTW) {
            // the exact first row, without validity guards in warps whose 96 candidates are all valid (7.17 -> 7.06 ms per pair at level 0)
            if (warp_all) refine_row_loop<STRIDE, false, Lut0, true>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, Lut0(), cs, ws, s_tile + pl, -PATCH_R, -PATCH_R);
            else refine_row_loop<STRIDE, true, Lut0, true>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, Lut0(), cs, ws, s_tile + pl, -PATCH_R, -PATCH_R);
            float lo = FLT_MAX;   // smallest accumulator among the valid candidates of this lane after the first row
#pragma unroll
            for (int m = 0; m < 3; m++)
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (valid[m]) lo = fminf(lo, fminf(cs[m][q], ws[m][q]));
            if (__all_sync(any_mask, lo >= 1.57772181e-30f)) {   // 2^-99; any_mask: lanes outside this branch (image border, unknown flow) do not vote
                if (warp_all) refine_row_loop<STRIDE, false, Lut0, true, false>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, Lut0(), cs, ws, s_tile + pl, -PATCH_R + STRIDE, PATCH_R);
                else refine_row_loop<STRIDE, true, Lut0, true, false>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, Lut0(), cs, ws, s_tile + pl, -PATCH_R + STRIDE, PATCH_R);
            } else
                refine_row_loop<STRIDE, true, Lut0, true>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, Lut0(), cs, ws, s_tile + pl, -PATCH_R + STRIDE, PATCH_R);
        } else if (TMA1) {
            refine_row_loop<STRIDE, true, Lut0, true>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, Lut0(), cs, ws, s_tile + pl);
        } else if (UNI) {
            refine_row_loop<STRIDE, false>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, Lut0(), cs, ws);
        } else if (LUTX) {
            const LutX<LUTX ? LUTX : 1> lx = {(unsigned)(pl % (LUTX ? LUTX : 1)) * 4u};
            refine_row_loop<STRIDE, true>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, lx, cs, ws);
        } else if (LUT0) {
            if (warp_all) refine_row_loop<STRIDE, false>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, Lut0(), cs, ws);
            else refine_row_loop<STRIDE, true>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, Lut0(), cs, ws);
        } else {
            if (warp_all) refine_row_loop<STRIDE, false>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, lut_base, cs, ws);
            else refine_row_loop<STRIDE, true>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, lut_base, cs, ws);
        }
#pragma unroll
        for (int m = 0; m < 3; m++) {
            if (!valid[m]) continue;
            const float k1 = __fdiv_rn(cs[m][0], ws[m][0]), k2 = __fdiv_rn(cs[m][1], ws[m][1]);
            const float k3 = __fdiv_rn(cs[m][2], ws[m][2]), k4 = __fdiv_rn(cs[m][3], ws[m][3]);
            cost[m] = min_ref(k1, min_ref(k2, min_ref(k3, k4)));  // :512
        }
    }
#pragma unroll
    for (int m = 0; m < 3; m++) s_best[m * 3 + n][pl] = valid[m] ? cost[m] : __int_as_float(0x7f800000);   // index = m * 3 + n: the reference's order (m outer)
    __syncthreads();
    if (in && n == 0) {
        float2 out;
        if (unknown) out = make_float2(0.f, 0.f);
        else {
            float bcost = 999999.f;   // :2024,:2031 strict '<' against 999999, m outer, n inner
            int bk = -1;
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const float oc = s_best[k][pl];
                if (oc < bcost) { bcost = oc; bk = k; }
            }
            short bx = cxc, by = cyc;   // :2020-2022 default = centre candidate
            if (bk >= 0) { bx = (short)(cxc + (bk / 3 - 1)); by = (short)(cyc + (bk % 3 - 1)); }
            out = make_float2((float)(bx - x), (float)(by - y));  // :2038-2039
        }
        a.flow[(size_t)b * a.w * a.h + (size_t)y * a.w + x] = out;
    }
}

// the shared-window address of the first byte of dynamic shared memory in a kernel without static shared memory (Lut0 relies on it)
__global__ void k_probe_shared_base(unsigned* out) {
    extern __shared__ float s_dyn[];
    s_dyn[threadIdx.x] = 0.f;
    if (threadIdx.x == 0) *out = (unsigned)__cvta_generic_to_shared(s_dyn);
}
static bool lut0_window_base_ok(int device) {
    static int cached[64] = {};   // 0 unknown, 1 ok, 2 not
    if (device >= 0 && device < 64 && cached[device]) return cached[device] == 1;
    unsigned* d = nullptr;
    unsigned h = ~0u;
    if (cudaMalloc((void**)&d, 4) == cudaSuccess) {
        k_probe_shared_base<<<1, 32, 256>>>(d);
        if (cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost) != cudaSuccess) h = ~0u;
        cudaFree(d);
    }
    const bool ok = h == SHARED_WINDOW_USER_BASE;
    if (device >= 0 && device < 64) cached[device] = ok ? 1 : 2;
    return ok;
}

// ---- table-driven refine with a SHARED AD + census volume (default at stride 2) ----
// k_c2f_refine_row scores 36 (candidate, model) pairs x 100 samples per pixel, each sample = an AD + census term (13 of its ~24 issued
// instructions, the POPC and one of its two MUFU.EX2) times a bilateral weight.  The AD + census term depends only on the image-1 pixel
// q1 = x + j of patch row i and on the displacement between the two sampled pixels, (dX, dY) = D(x) + (m - 1, n - 1) + site offset of the
// model -- not on the pixel x that asks for it: along a patch row, pixel x at sample j + 2 and pixel x + 2 at sample j ask for the same term
// whenever their integer flows agree.  Per patch row the CTA therefore first fills a volume V[line(dX, dY)][q1] for its 50 image-1 columns
// and the box of displacements its 32 pixels x 9 candidates x 4 models can ask for (VolTab; 21-48 of the box's lines are ever read, against
// 360 direct evaluations per column), coalesced and with the very instruction sequence of sample_eval, and the scoring loop reads the term
// back with one LDS per sample.  The weight half of a sample -- which depends on the candidate's centre pixel -- is computed as before, in
// the reference's sample order: the same bits.  CTAs whose integer flows spread by more than 1, that touch the image border, or that
// hold an unknown flow keep the direct loop of k_c2f_refine_row.
// MEASURED SLOWER than k_c2f_refine_row and therefore not the default (EPPM_VARIANT bit 8388608 / EPPM_REFINE_MODE 20-22; same bits): 8.75 ms per
// 1080p pair at level 0 against 7.54.  The scoring loop drops from ~25 to 15.2 issued instructions per sample and the XU pipe from 63 % to 34 %,
// but (a) filling the volume costs ~80 instructions per line (342-495 lines per CTA, 50 of 64 lane slots used), which brings the total back
// to the direct kernel's count (24.8e9 against 25.5e9 warp instructions per 4 pairs), (b) the 30 KB of shared memory per CTA leave 6 CTAs per
// SM and 42 KB of L1 (7 CTAs / 11 KB of L1: 12.3 ms), and (c) both forms need the same five L1 wavefronts per sample -- four for the 16-byte
// image-2 load the weight needs, one for the census table or the volume -- and with fewer instructions the L1 data pipe becomes the limit
// (82 % of its peak, 60 % of the issue slots): profiles/r02_ncu_refine_vol_l0_summary.txt.
__device__ __forceinline__ float ad_census_term(const float4& p1, const float4& p2) {
    const float c = max3abs_diff(pack_pix(p1), pack_pix(p2));
    return __fadd_rn(exp_ad_cost(c), census_lut_ref(Lut0(), p1, p2));   // (1 - e) + census: sample_eval's two roundings
}
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
template <int IMM>
__device__ __forceinline__ float lds_f32_imm(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    return v;
}

// one patch row of the scoring loop: weights as in refine_row_loop, AD + census from the volume.  lane_row: shared-window byte address of
// this lane's column in line ((n - 1 + ddy) * bx + ddx) of the volume.
template <bool TINY>
__device__ __forceinline__ void vol_use_row(const RefineArgs& a, const CostLut& lut, const AffineTab& tab, const VolTab& vt, int r, const float4* trow,
                                            const float4* Pc, const PixPk& c1k, const PixPk (&c2k)[3], unsigned wmask, unsigned lane_row,
                                            float (&cs)[3][4], float (&ws)[3][4]) {
    const int i = -PATCH_R + 2 * r;
    const int ai = i < 0 ? -i : i;
    const int irow = i * a.pw;
    int s = r * 10;
EPPM_PRAGMA(unroll RF_JUNROLL)
    for (int j = -PATCH_R; j <= PATCH_R; j += 2, s++) {
        const PixPk p1k = pack_pix(trow[j + PATCH_R]);
        const float d1 = max3abs_diff(c1k, p1k);
        const float gg = lut.gg[ai][j < 0 ? -j : j];
        const float4* site[4];
        unsigned va[4];
        site[0] = pix_at(Pc, irow + j);
        va[0] = lane_row + (unsigned)vt.T[0][s];
#pragma unroll
        for (int q = 1; q < 4; q++) {
            site[q] = pix_at(Pc, tab.off[q - 1][s]);
            va[q] = lane_row + (unsigned)vt.T[q][s];
        }
#pragma unroll
        for (int m = 0; m < 3; m++) {
            if (!(wmask & (1u << m))) continue;   // uniform, never taken: keeps the three candidates in separate basic blocks (see refine_row_loop)
            float arg[4], t2[4], w[4], ct[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float d2 = max3abs_diff(c2k[m], pack_pix(ldpix(site[q] + (m - 1))));
                arg[q] = __fmaf_rn(d1, d1, __fmul_rn(d2, d2));
            }
            const f32x2 R = pk2(-99.99999237060546875f, -99.99999237060546875f), D = pk2(0.010000000707805156708f, 0.010000000707805156708f);
            const f32x2 Z = pk2(0.f, 0.f), L2E = pk2(1.4426950216293334961f, 1.4426950216293334961f);
#pragma unroll
            for (int h = 0; h < 2; h++) {   // div_neg_0p01 and the log2e multiply on two models per packed instruction (IEEE per half)
                const f32x2 x = pk2(arg[2 * h], arg[2 * h + 1]);
                const f32x2 q0 = fma2(x, R, Z);
                const f32x2 qq = fma2(R, fma2(q0, D, x), q0);
                upk2(mul2(qq, L2E), t2[2 * h], t2[2 * h + 1]);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) w[q] = __fmul_rn(ex2_mufu(t2[q]), gg);
            if (TINY && fminf(fminf(t2[0], t2[1]), fminf(t2[2], t2[3])) < -126.0f) {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (t2[q] < -126.0f) w[q] = __fmul_rn(ex2_tiny(t2[q]), gg);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                ct[q] = m == 0 ? lds_f32_imm<-4 * VOL_COLS>(va[q]) : m == 1 ? lds_f32_imm<0>(va[q]) : lds_f32_imm<4 * VOL_COLS>(va[q]);
                cs[m][q] = __fmaf_rn(ct[q], w[q], cs[m][q]);
                ws[m][q] = __fadd_rn(ws[m][q], w[q]);
            }
        }
    }
}

constexpr int RF_VOL_TILE_OFF = 1280;                                                              // byte offsets inside the dynamic shared memory
constexpr int RF_VOL_BAR_OFF = RF_VOL_TILE_OFF + RF_TILE_W * RF_TILE_H * (int)sizeof(float4);
constexpr int RF_VOL_V_OFF = RF_VOL_BAR_OFF + 16;
constexpr int RF_VOL_SMEM = RF_VOL_V_OFF + VOL_MAX_LINES * VOL_COLS * (int)sizeof(float);

template <int MINB>
__global__ void __launch_bounds__(RF_PIX * 3, MINB)
    k_c2f_refine_vol(RefineArgs a, const __grid_constant__ CostLut lut, const __grid_constant__ AffineTab tab, const __grid_constant__ CUtensorMap tmap1,
                     const __grid_constant__ VolTab vt) {
    // no static shared memory in this kernel (Lut0): [0, 16) census table, s_best[9][RF_PIX], image-1 tile, mbarrier, volume
    extern __shared__ float s_dyn[];
    float (*s_best)[RF_PIX] = reinterpret_cast<float (*)[RF_PIX]>(s_dyn + 16);
    float4* s_tile = reinterpret_cast<float4*>(reinterpret_cast<char*>(s_dyn) + RF_VOL_TILE_OFF);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(s_dyn) + RF_VOL_BAR_OFF);
    float* s_vol = reinterpret_cast<float*>(reinterpret_cast<char*>(s_dyn) + RF_VOL_V_OFF);
    load_census_lut(s_dyn, lut);
    const int n = threadIdx.x >> 5, pl = threadIdx.x & 31;   // n: candidate row of this warp
    const int x0 = blockIdx.x * RF_PIX, x = x0 + pl, y = a.y0 + blockIdx.y;
    const bool in = x < a.w;
    const int b = blockIdx.z;
    if (threadIdx.x == 0) {
        mbar_init(s_bar, 1);
        mbar_expect_tx(s_bar, RF_TILE_W * RF_TILE_H * sizeof(float4));
        tma_load_3d(s_tile, &tmap1, (x0 + PAD - PATCH_R) * 4, y + PAD - PATCH_R, b, s_bar);
    }
    const float4* I1 = a.pix1 + (size_t)b * a.plane;
    const float4* I2 = a.pix2 + (size_t)b * a.plane;
    float2 fl = make_float2(0.f, 0.f);
    if (in) fl = a.upsample ? upsample2(a.coarse + (size_t)b * a.ws * a.hs, a.ws, a.hs, x, y) : a.coarse[(size_t)b * a.w * a.h + (size_t)y * a.w + x];
    const bool unknown = fl.x > EPPM_UNKNOWN_FLOW_THRESH || fl.y > EPPM_UNKNOWN_FLOW_THRESH;  // :2011
    const short cxc = (short)((short)(int)fl.x + x), cyc = (short)((short)(int)fl.y + y);     // :2014-2019
    const short cy = (short)(cyc + (n - 1));
    __syncthreads();          // the barrier object is initialised before anyone polls it
    mbar_wait(s_bar, 0);
    float cost[3];
    bool valid[3];
    bool any = false, all = true;
#pragma unroll
    for (int m = 0; m < 3; m++) {
        const short cx = (short)(cxc + (m - 1));
        cost[m] = FLT_MAX;
        valid[m] = in && !unknown && !(cx < 0 || cy < 0 || cx >= a.w || cy >= a.h);  // :2029
        any = any || valid[m];
        all = all && valid[m];
    }
    // CTA-uniform choice (the three warps hold the same 32 pixels, so each of them computes the same answer): the volume needs every lane's nine
    // candidates inside the image -- then no site leaves the padded plane, see build_affine_tab -- and integer flows within one step of each other
    const int dxl = (int)cxc - x, dyl = (int)cyc - y;
    const bool inside = in && !unknown && cxc >= 1 && cxc + 1 < a.w && cyc >= 1 && cyc + 1 < a.h;
    const bool full = __all_sync(0xffffffffu, inside);
    const int Dx0 = __reduce_min_sync(0xffffffffu, dxl), Dx1 = __reduce_max_sync(0xffffffffu, dxl);
    const int Dy0 = __reduce_min_sync(0xffffffffu, dyl), Dy1 = __reduce_max_sync(0xffffffffu, dyl);
    const bool use_vol = full && Dx1 - Dx0 <= 1 && Dy1 - Dy0 <= 1;
    unsigned wmask = 7u;
    asm volatile("" : "+r"(wmask));   // opaque to the compiler: see refine_row_loop
    float cs[3][4], ws[3][4];
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
        for (int q = 0; q < 4; q++) cs[m][q] = ws[m][q] = 0.f;
    if (use_vol) {
        const PixPk c1k = pack_pix(ldpix(I1 + (unsigned)((y + PAD) * a.pw + x + PAD)));
        const float4* Pc = I2 + (unsigned)(((int)cy + PAD) * a.pw + (int)cxc + PAD);
        PixPk c2k[3];
#pragma unroll
        for (int m = 0; m < 3; m++) c2k[m] = pack_pix(ldpix(Pc + (m - 1)));
        const int sxy = (Dx1 - Dx0) | ((Dy1 - Dy0) << 1);
        const unsigned vs = (unsigned)__cvta_generic_to_shared(s_vol);
        bool fast = false;
#pragma unroll 1
        for (int r = 0; r < 10; r++) {
            const int bxr = vt.bx[r], nl = bxr * vt.by[r];
            {   // fill the lines of this patch row: warp n takes lines n, n + 3, ...; a lane its column and, for lanes < 18, column + 32
                const int i = -PATCH_R + 2 * r;
                const unsigned rowbase = (unsigned)((y + i + Dy0 + vt.ylo[r] + PAD) * a.pw + (x0 - PATCH_R + Dx0 + vt.xlo[r] + PAD));
                const float4* trow = s_tile + r * RF_TILE_W;
                int dXi = n, dYi = 0;
#pragma unroll 1
                for (int L = n; L < nl; L += 3) {
                    if ((vt.used[r][sxy][L >> 5] >> (L & 31)) & 1u) {
                        const float4* rp = I2 + (rowbase + (unsigned)(dYi * a.pw + dXi));
                        float* vrow = s_vol + L * VOL_COLS;
                        const float4 p2a = ldpix(rp + pl);
                        float4 p2b = p2a;
                        if (pl < VOL_COLS - 32) p2b = ldpix(rp + pl + 32);
                        vrow[pl] = ad_census_term(trow[pl], p2a);
                        if (pl < VOL_COLS - 32) vrow[pl + 32] = ad_census_term(trow[pl + 32], p2b);
                    }
                    dXi += 3;
                    if (dXi >= bxr) { dXi -= bxr; dYi++; }
                }
            }
            __syncthreads();
            const unsigned lane_row = vs + 4u * (unsigned)(((n - 1 + (dyl - Dy0)) * bxr + (dxl - Dx0)) * VOL_COLS + pl);
            const float4* trow = s_tile + r * RF_TILE_W + pl;
            if (fast) vol_use_row<false>(a, lut, tab, vt, r, trow, Pc, c1k, c2k, wmask, lane_row, cs, ws);
            else vol_use_row<true>(a, lut, tab, vt, r, trow, Pc, c1k, c2k, wmask, lane_row, cs, ws);
            if (r == 0) {   // FASTW (see k_c2f_refine_row): every accumulator >= 2^-99 absorbs any weight that would need the __expf fix-up
                float lo = FLT_MAX;
#pragma unroll
                for (int m = 0; m < 3; m++)
#pragma unroll
                    for (int q = 0; q < 4; q++) lo = fminf(lo, fminf(cs[m][q], ws[m][q]));
                fast = __all_sync(0xffffffffu, lo >= 1.57772181e-30f);
            }
            __syncthreads();   // the next patch row overwrites the volume
        }
    } else if (any) {
        const float4* a0 = I1 + (unsigned)((y + PAD) * a.pw + x + PAD);
        const PixPk c1k = pack_pix(ldpix(a0));
        const int cys = max(0, min(a.h - 1, (int)cy)), cxs = max(-1, min(a.w, (int)cxc));
        const float4* Pc = I2 + (unsigned)((cys + PAD) * a.pw + cxs + PAD);
        asm volatile("" : "+l"(Pc));
        PixPk c2k[3];
#pragma unroll
        for (int m = 0; m < 3; m++) c2k[m] = pack_pix(ldpix(Pc + (m - 1)));
        refine_row_loop<2, true, Lut0, true>(a, lut, tab, a0, Pc, c1k, c2k, valid, wmask, Lut0(), cs, ws, s_tile + pl);
    }
    if (any) {
#pragma unroll
        for (int m = 0; m < 3; m++) {
            if (!valid[m]) continue;
            const float k1 = __fdiv_rn(cs[m][0], ws[m][0]), k2 = __fdiv_rn(cs[m][1], ws[m][1]);
            const float k3 = __fdiv_rn(cs[m][2], ws[m][2]), k4 = __fdiv_rn(cs[m][3], ws[m][3]);
            cost[m] = min_ref(k1, min_ref(k2, min_ref(k3, k4)));  // :512
        }
    }
#pragma unroll
    for (int m = 0; m < 3; m++) s_best[m * 3 + n][pl] = valid[m] ? cost[m] : __int_as_float(0x7f800000);   // index = m * 3 + n: the reference's order (m outer)
    __syncthreads();
    if (in && n == 0) {
        float2 out;
        if (unknown) out = make_float2(0.f, 0.f);
        else {
            float bcost = 999999.f;   // :2024,:2031 strict '<' against 999999, m outer, n inner
            int bk = -1;
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const float oc = s_best[k][pl];
                if (oc < bcost) { bcost = oc; bk = k; }
            }
            short bx = cxc, by = cyc;   // :2020-2022 default = centre candidate
            if (bk >= 0) { bx = (short)(cxc + (bk / 3 - 1)); by = (short)(cyc + (bk % 3 - 1)); }
            out = make_float2((float)(bx - x), (float)(by - y));  // :2038-2039
        }
        a.flow[(size_t)b * a.w * a.h + (size_t)y * a.w + x] = out;
    }
}

// ---- table-driven refine, models in packed pairs (default) ----
// Same decomposition as k_c2f_refine_tab with NCT = 3 (warp m owns candidate column m, a thread its three candidate rows x four
// models), but the four models of a candidate are evaluated as TWO PAIRS in the halves of packed FP32x2 instructions
// (sample_eval2): squares, both constant divisions, log2e, 1 - e, + census, e * gg and the two accumulations cost one issue slot per
// pair instead of one per model.  Accumulators are pairs too: cs[n][0] = (model 0, model 1), cs[n][1] = (model 2, model 3).  Each
// (candidate, model) still adds its samples in the reference's order with per-half IEEE rounding: the same bits.
// ALLROWS: candidate rows that are not valid (outside the image, :2029) are scored at a clamped centre and discarded instead of
// being branched around -- only warps at the image border contain such rows, and the branch costs three issue slots per candidate
// and sample everywhere else.
#ifndef RF_PK_MINBLOCKS
#define RF_PK_MINBLOCKS 7
#endif
#ifndef RF_PK_LUT
#define RF_PK_LUT lut_base
#endif
#ifdef RF_PK_MAXNREG
#define RF_PK_BOUNDS __maxnreg__(RF_PK_MAXNREG)
#else
#define RF_PK_BOUNDS __launch_bounds__(RF_PIX * 3, MINB)
#endif
template <int MINB, int STRIDE, bool ALLROWS>
__global__ void RF_PK_BOUNDS
    k_c2f_refine_pk(RefineArgs a, const __grid_constant__ CostLut lut, const __grid_constant__ AffineTab tab) {
    __shared__ float s_best[9][RF_PIX];
    __shared__ float s_census[CENSUS_LUT_N];
    load_census_lut(s_census, lut);
    const unsigned lut_base = census_lut_base(s_census);
    const int m = threadIdx.x >> 5, pl = threadIdx.x & 31;
    const int x = blockIdx.x * RF_PIX + pl, y = a.y0 + blockIdx.y;
    const bool in = x < a.w;
    const int b = blockIdx.z;
    const float4* I1 = a.pix1 + (size_t)b * a.plane;
    const float4* I2 = a.pix2 + (size_t)b * a.plane;
    float2 fl = make_float2(0.f, 0.f);
    if (in) fl = a.upsample ? upsample2(a.coarse + (size_t)b * a.ws * a.hs, a.ws, a.hs, x, y) : a.coarse[(size_t)b * a.w * a.h + (size_t)y * a.w + x];
    const bool unknown = fl.x > EPPM_UNKNOWN_FLOW_THRESH || fl.y > EPPM_UNKNOWN_FLOW_THRESH;  // :2011
    const short cxc = (short)((short)(int)fl.x + x), cyc = (short)((short)(int)fl.y + y);     // :2014-2019
    const short cx = (short)(cxc + (m - 1));
    float cost[3];
    bool valid[3];
    bool any = false;
#pragma unroll
    for (int n = 0; n < 3; n++) {
        const short cy = (short)(cyc + (n - 1));
        cost[n] = FLT_MAX;
        valid[n] = in && !unknown && !(cx < 0 || cy < 0 || cx >= a.w || cy >= a.h);  // :2029
        any = any || valid[n];
    }
    if (any) {
        f32x2 cs[3][2], ws[3][2];
#pragma unroll
        for (int n = 0; n < 3; n++)
#pragma unroll
            for (int q = 0; q < 2; q++) cs[n][q] = ws[n][q] = pk2(0.f, 0.f);
        const float4* a0 = I1 + (unsigned)((y + PAD) * a.pw + x + PAD);
        const PixPk c1k = pack_pix(ldpix(a0));
        PixPk c2k[3];
        const float4* P[3];
#pragma unroll
        for (int n = 0; n < 3; n++) {
            const int cy = max(0, min(a.h - 1, (int)cyc + n - 1));   // rows that are not valid are scored at a clamped centre and never used
            const int cxs = max(0, min(a.w - 1, (int)cx));
            P[n] = I2 + (unsigned)((cy + PAD) * a.pw + cxs + PAD);
            c2k[n] = pack_pix(ldpix(P[n]));
            asm volatile("" : "+l"(P[n]));
        }
        int s = 0;
#pragma unroll 1
        for (int i = -PATCH_R; i <= PATCH_R; i += STRIDE) {
            const int ai = i < 0 ? -i : i;
            const int irow = i * a.pw;
EPPM_PRAGMA(unroll RF_JUNROLL)
            for (int j = -PATCH_R; j <= PATCH_R; j += STRIDE, s++) {
                const float4 p1 = ldpix(a0 + irow + j);
                const PixPk p1k = pack_pix(p1);
                const float d1s = max3abs_diff(c1k, p1k);
                const f32x2 d1 = pk2(d1s, d1s);
                const float ggs = lut.gg[ai][j < 0 ? -j : j];
                int off[4];
                off[0] = irow + j;  // identity model: the exact integer site (cx + j, cy + i)
#pragma unroll
                for (int q = 0; q < 3; q++) off[q + 1] = tab.off[q][s];
                if (STRIDE == 1 && s == tab.exc_s) {   // the one site of stride 1 whose x offset depends on the coordinate itself (see AffineTab)
                    const int X = (int)cx + j;
                    const int dx = __float2int_rd(__fmaf_rn((float)i, tab.exc_ci, __fmaf_rn((float)j, tab.exc_cj, (float)X))) - X;
#pragma unroll
                    for (int q = 0; q < 3; q++)
                        if (q == tab.exc_q) off[q + 1] += dx;
                }
#pragma unroll
                for (int n = 0; n < 3; n++) {
                    if (!ALLROWS && !valid[n]) continue;
                    const int pxb = a.px_bytes;
                    f32x2 ct[2], t2[2], w[2];
#pragma unroll
                    for (int h = 0; h < 2; h++)
                        sample_eval2(p1, p1k, p1, p1k, ldpix(pix_at_b(P[n], off[2 * h], pxb)), ldpix(pix_at_b(P[n], off[2 * h + 1], pxb)), c2k[n], c2k[n], d1,
                                     RF_PK_LUT, ct[h], t2[h]);
                    sample_weight4(t2[0], t2[1], ggs, w[0], w[1]);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        cs[n][h] = fma2(ct[h], w[h], cs[n][h]);
                        ws[n][h] = add2(ws[n][h], w[h]);
                    }
                }
            }
        }
#pragma unroll
        for (int n = 0; n < 3; n++) {
            if (!valid[n]) continue;
            float c0, c1, c2, c3, w0, w1, w2, w3;
            upk2(cs[n][0], c0, c1); upk2(cs[n][1], c2, c3);
            upk2(ws[n][0], w0, w1); upk2(ws[n][1], w2, w3);
            const float k1 = __fdiv_rn(c0, w0), k2 = __fdiv_rn(c1, w1), k3 = __fdiv_rn(c2, w2), k4 = __fdiv_rn(c3, w3);
            cost[n] = min_ref(k1, min_ref(k2, min_ref(k3, k4)));  // :512
        }
    }
    // candidates that are not valid never win: the reference skips them (:2029), here they carry +inf against the strict '<' below
#pragma unroll
    for (int n = 0; n < 3; n++) s_best[m * 3 + n][pl] = valid[n] ? cost[n] : __int_as_float(0x7f800000);
    __syncthreads();
    if (in && m == 0) {
        float2 out;
        if (unknown) out = make_float2(0.f, 0.f);
        else {
            // arg-min in the reference's order: m outer, n inner, strict '<' against 999999 (:2024,:2031)
            float bcost = 999999.f;
            int bk = -1;
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const float oc = s_best[k][pl];
                if (oc < bcost) { bcost = oc; bk = k; }
            }
            short bx = cxc, by = cyc;   // :2020-2022 default = centre candidate
            if (bk >= 0) { bx = (short)(cxc + (bk / 3 - 1)); by = (short)(cyc + (bk % 3 - 1)); }
            out = make_float2((float)(bx - x), (float)(by - y));  // :2038-2039
        }
        a.flow[(size_t)b * a.w * a.h + (size_t)y * a.w + x] = out;
    }
}

// Site table of one level (pitch pw) and its proof: for every integer X in [lo, hi] that a candidate coordinate + offset can take,
// floor(fmaf(i, Cy, fmaf(j, Cx, (float)X))) - X must equal the tabulated value.  Host fmaf is correctly rounded = the device FFMA.
bool build_affine_tab(AffineTab& t, int pw, int w, int h, int stride, bool allow_exception) {
    static const float pf[3][4] = {{0.177f, -0.011f, -0.003f, 0.301f}, {0.125f, -0.357f, 0.009f, 0.308f}, {0.205f, 0.370f, 0.011f, 0.296f}};
    auto site = [](float fi, float fj, float cj, float ci, int X) { return (int)floorf(fmaf(fi, ci, fmaf(fj, cj, (float)X))) - X; };
    if (stride < 1 || stride > 3) return false;
    const int lim = (w > h ? w : h) + PATCH_R;
    t.exc_s = -1; t.exc_q = 0; t.exc_cj = t.exc_ci = 0.f;
    int s = 0;
    for (int i = -PATCH_R; i <= PATCH_R; i += stride)
        for (int j = -PATCH_R; j <= PATCH_R; j += stride, s++)
            for (int q = 0; q < 3; q++) {
                // x: cx2 = fma(i, C_uy, fma(j, C_ux, float(cx + j)));  y: cy2 = fma(i, C_vy, fma(j, C_vx, float(cy + i)))
                int dx = site((float)i, (float)j, pf[q][0], pf[q][1], 1000);
                const int dy = site((float)i, (float)j, pf[q][2], pf[q][3], 1000);
                bool x_ok = true;
                for (int X = -PATCH_R; X <= lim; X++) {
                    if (site((float)i, (float)j, pf[q][2], pf[q][3], X) != dy) return false;
                    const int d = site((float)i, (float)j, pf[q][0], pf[q][1], X);
                    if (d != dx) x_ok = false;
                    if (abs(d + j) >= PAD) return false;
                }
                if (!x_ok) {   // the x offset of this site depends on X: one such site may be left to the kernel
                    if (!allow_exception || t.exc_s >= 0) return false;
                    t.exc_s = s; t.exc_q = q; t.exc_cj = pf[q][0]; t.exc_ci = pf[q][1];
                    dx = 0;
                }
                if (abs(dy + i) >= PAD) return false;
                t.off[q][s] = (dy + i) * pw + (dx + j);
            }
    return true;
}

// Tables of the shared AD + census volume (see VolTab).  Site offsets as in build_affine_tab (whose check that they do not depend on the
// coordinate is what makes a table legitimate; the volume kernel is only used at levels whose affine table verified).
bool build_vol_tab(VolTab& v) {
    static const float pf[3][4] = {{0.177f, -0.011f, -0.003f, 0.301f}, {0.125f, -0.357f, 0.009f, 0.308f}, {0.205f, 0.370f, 0.011f, 0.296f}};
    auto site = [](float fi, float fj, float cj, float ci, int X) { return (int)floorf(fmaf(fi, ci, fmaf(fj, cj, (float)X))) - X; };
    static int ox[4][100], oy[4][100];
    int s = 0;
    for (int i = -PATCH_R; i <= PATCH_R; i += 2)
        for (int j = -PATCH_R; j <= PATCH_R; j += 2, s++) {
            ox[0][s] = oy[0][s] = 0;
            for (int q = 0; q < 3; q++) {
                ox[q + 1][s] = site((float)i, (float)j, pf[q][0], pf[q][1], 1000);
                oy[q + 1][s] = site((float)i, (float)j, pf[q][2], pf[q][3], 1000);
            }
        }
    memset(&v, 0, sizeof(v));
    for (int r = 0; r < 10; r++) {
        int xmin = 1 << 20, xmax = -(1 << 20), ymin = 1 << 20, ymax = -(1 << 20);
        for (int q = 0; q < 4; q++)
            for (int jj = 0; jj < 10; jj++) {
                const int k = r * 10 + jj;
                xmin = ox[q][k] < xmin ? ox[q][k] : xmin; xmax = ox[q][k] > xmax ? ox[q][k] : xmax;
                ymin = oy[q][k] < ymin ? oy[q][k] : ymin; ymax = oy[q][k] > ymax ? oy[q][k] : ymax;
            }
        // box = site offsets + candidate column / row (-1..1) + flow spread inside the CTA (0..1)
        v.xlo[r] = xmin - 1; v.bx[r] = xmax - xmin + 4;
        v.ylo[r] = ymin - 1; v.by[r] = ymax - ymin + 4;
        if (v.bx[r] < 4 || v.bx[r] * v.by[r] > VOL_MAX_LINES || v.bx[r] * v.by[r] > 128) return false;
        for (int q = 0; q < 4; q++)
            for (int jj = 0; jj < 10; jj++) {
                const int k = r * 10 + jj;
                const int line = (oy[q][k] - v.ylo[r]) * v.bx[r] + (ox[q][k] - v.xlo[r]);
                v.T[q][k] = 4 * (line * VOL_COLS + 2 * jj);   // column of lane 0 at sample j = -9 + 2 jj: (j + 9)
                for (int sxy = 0; sxy < 4; sxy++)
                    for (int m = -1; m <= 1; m++)
                        for (int n = -1; n <= 1; n++)
                            for (int ddx = 0; ddx <= (sxy & 1); ddx++)
                                for (int ddy = 0; ddy <= (sxy >> 1); ddy++) {
                                    const int L = line + (n + ddy) * v.bx[r] + m + ddx;
                                    v.used[r][sxy][L >> 5] |= 1u << (L & 31);
                                }
            }
    }
    return true;
}

// d_flow_bilateral_filtering (bao_pmflow_refine_kernel.cu:764-799): (2R+1)^2 joint bilateral, R = 2*sig_s, skipping unknown
// flow, snapshot semantics.  Tile of flow + image-1 colours with R halo in shared memory.
struct SmoothArgs {
    const float2* src;
    float2* dst;
    const float4* pix;  // padded origin, image 1
    size_t plane;
    int pw, w, h;
    int R;
    float neg_sig_r2;   // -(sig_r*sig_r), the divisor nvcc folds `-(d*d)/(SIG_R*SIG_R)` into
    float recip;        // RN(1/neg_sig_r2)
    int y0, y1;         // row band [y0, y1) this launch owns
    int fast_div;       // 1: x/neg_sig_r2 as q0=x*r, rem=fma(-q0,d,x), q=fma(rem,r,q0) (exactness verified by eppm_selftest_const_div)
};

// CTA = 32 x 8 threads, every thread filters TWO vertically adjacent pixels (y, y+1): a tile row is tap row dy of the upper
// pixel and dy-1 of the lower one, so each shared-memory read (16 B colour + 8 B flow) feeds two taps -- the kernel was
// shared-memory-bandwidth bound with one pixel per thread.  Each pixel still accumulates its taps in (dy, dx) raster order.
constexpr int SM_TX = 32, SM_TY = 8, SM_PY = 2;
__device__ __forceinline__ void smooth_tap(const SmoothArgs& a, const float4& c, const float4& p, const float2& fl, float gg, float r, float nd, float& nx,
                                           float& ny, float& wsum) {
    const float dr = max3abs_diff(p, c);                                              // :757
    const float xx = __fmul_rn(dr, dr);
    float q;
    if (a.fast_div) {
        const float q0 = __fmul_rn(xx, r);
        q = __fmaf_rn(__fmaf_rn(q0, nd, xx), r, q0);
    } else {
        q = __fdiv_rn(xx, a.neg_sig_r2);
    }
    const float wgt = __fmul_rn(exp_ref(q), gg);                                      // :758-760
    nx = __fmaf_rn(wgt, fl.x, nx);                                                    // :782-783
    ny = __fmaf_rn(wgt, fl.y, ny);
    wsum = __fadd_rn(wsum, wgt);
}

__global__ void __launch_bounds__(SM_TX* SM_TY) k_flow_smooth(SmoothArgs a, const __grid_constant__ SmoothLut lut, const __grid_constant__ CUtensorMap tmap,
                                                              int use_tma) {
    extern __shared__ __align__(128) float4 smem[];
    __shared__ __align__(8) unsigned long long s_bar;  // [TH*TW] colours, then float2 [TH*TW] flows, then float [(R+1)^2] spatial weights
    const int R = a.R, TW = SM_TX + 2 * R, TH = SM_TY * SM_PY + 2 * R;
    float4* s_pix = smem;
    float2* s_flow = reinterpret_cast<float2*>(smem + TW * TH);
    float* s_gg = reinterpret_cast<float*>(s_flow + TW * TH);
    const int b = blockIdx.z;
    const float2* f = a.src + (size_t)b * a.w * a.h;
    const float4* img = a.pix + (size_t)b * a.plane + (size_t)PAD * a.pw + PAD;
    const int x0 = blockIdx.x * SM_TX - R, y0 = a.y0 + blockIdx.y * SM_TY * SM_PY - R;
    const int tid = threadIdx.y * SM_TX + threadIdx.x;
    for (int i = tid; i < (R + 1) * (R + 1); i += SM_TX * SM_TY)
        s_gg[i] = __fmul_rn(lut.g[i % (R + 1)], lut.g[i / (R + 1)]);  // cBlfGaussian[|dx|] * cBlfGaussian[|dy|] (:759)
    // Colour tile (+R halo) of image 1: ONE TMA bulk tensor copy from the padded packed plane (box TW x TH pixels of 16 B; coordinates in
    // floats along x; out-of-range parts are zero-filled and never used because their flow is marked unknown below).
    if (use_tma) {
        if (tid == 0) {
            mbar_init(&s_bar, 1);
        }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&s_bar, (unsigned)(TW * TH * sizeof(float4)));
            tma_load_3d(s_pix, &tmap, (x0 + PAD) * 4, y0 + PAD, b, &s_bar);
        }
    }
    for (int i = tid; i < TW * TH; i += SM_TX * SM_TY) {
        const int ty = i / TW, tx = i % TW;
        const int cx = x0 + tx, cy = y0 + ty;
        float2 fl = make_float2(EPPM_UNKNOWN_FLOW, EPPM_UNKNOWN_FLOW);  // outside the image: skipped like unknown flow (:776,:778)
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cx >= 0 && cy >= 0 && cx < a.w && cy < a.h) {
            fl = f[(size_t)cy * a.w + cx];
            if (!use_tma) p = ldpix(img + (size_t)cy * a.pw + cx);
        }
        s_flow[i] = fl;
        if (!use_tma) s_pix[i] = p;
    }
    if (use_tma) mbar_wait(&s_bar, 0);
    __syncthreads();
    const int x = blockIdx.x * SM_TX + threadIdx.x;
    const int ly = threadIdx.y * SM_PY;               // local row of the upper pixel
    const int y = a.y0 + blockIdx.y * SM_TY * SM_PY + ly;
    if (x >= a.w || y >= a.y1) return;
    const float4 cA = s_pix[(ly + R) * TW + threadIdx.x + R];
    const float4 cB = s_pix[(ly + 1 + R) * TW + threadIdx.x + R];
    const float r = a.recip, nd = -a.neg_sig_r2;
    float nxA = 0.f, nyA = 0.f, wA = 0.f, nxB = 0.f, nyB = 0.f, wB = 0.f;
    // tile rows ly .. ly + 2R + 1: row t is tap dy = t - R of pixel A and dy = t - R - 1 of pixel B
    for (int t = 0; t <= 2 * R + 1; t++) {
        const int dyA = t - R, dyB = t - R - 1;
        const bool useA = dyA <= R, useB = dyB >= -R;
        const float* ggA = s_gg + abs(dyA) * (R + 1);
        const float* ggB = s_gg + abs(dyB) * (R + 1);
        const int rowb = (ly + t) * TW + threadIdx.x + R;
#pragma unroll 3
        for (int dx = -R; dx <= R; dx++) {
            const float2 fl = s_flow[rowb + dx];
            if (fmaxf(fl.x, fl.y) > EPPM_UNKNOWN_FLOW_THRESH) continue;                        // :778 (same truth table as x > T || y > T)
            const float4 p = s_pix[rowb + dx];
            const int adx = abs(dx);
            if (useA) smooth_tap(a, cA, p, fl, ggA[adx], r, nd, nxA, nyA, wA);
            if (useB) smooth_tap(a, cB, p, fl, ggB[adx], r, nd, nxB, nyB, wB);
        }
    }
    float2 outA = s_flow[(ly + R) * TW + threadIdx.x + R];
    if (wA != 0.f) outA = make_float2(__fdiv_rn(nxA, wA), __fdiv_rn(nyA, wA));  // :790-796 (untouched otherwise)
    a.dst[(size_t)b * a.w * a.h + (size_t)y * a.w + x] = outA;
    if (y + 1 < a.y1) {
        float2 outB = s_flow[(ly + 1 + R) * TW + threadIdx.x + R];
        if (wB != 0.f) outB = make_float2(__fdiv_rn(nxB, wB), __fdiv_rn(nyB, wB));
        a.dst[(size_t)b * a.w * a.h + (size_t)(y + 1) * a.w + x] = outB;
    }
}

// ---- smoothing, four rows per thread, packed pairs (default when R = 10) ----
// CTA = 32 x 8 threads, every thread filters FOUR vertically adjacent pixels (y .. y+3) as two pairs (A,B) and (C,D): tile row t is
// tap row dy = t - R - k of pixel k, so one shared-memory read of a tap (16 B colour + 8 B flow) feeds four pixels, and the two
// pixels of a pair run side by side in the halves of packed FP32x2 instructions (sub, mul, fma -- each half IEEE-rounded like the
// scalar instruction).  Every pixel still adds its 441 taps in (dy, dx) raster order.  Branch-free inner loop:
//  * taps outside the image or with unknown flow carry the colour SMOOTH_FAR in the tile: their range weight is exp(-huge) = +0
//    exactly and their flow is stored as 0, so fma(+0, 0, n) = n and w + 0 = w -- the reference's `continue` (:776,:778) bit for bit;
//  * tile rows that are out of a pixel's 21-row window (one row per pixel at either end of a pair) get a spatial weight of 0.
constexpr int S4_TX = 32, S4_TY = 8, S4_PY = 4, S4_R = 10;
constexpr int S4_TW = S4_TX + 2 * S4_R, S4_TH = S4_TY * S4_PY + 2 * S4_R, S4_ROWS = 2 * S4_R + S4_PY;   // 52 x 52 tile, 24 tile rows per thread
constexpr float SMOOTH_FAR = 1.0e4f;
constexpr size_t S4_SMEM = (size_t)S4_TW * S4_TH * (sizeof(float4) + sizeof(float2)) + 2 * S4_ROWS * (S4_R + 1) * sizeof(float2);

template <bool FAST_DIV>
__device__ __forceinline__ void smooth_tap2(const SmoothArgs& a, f32x2 cx, f32x2 cy, f32x2 cz, const float4& p, const float2& fl, f32x2 gg, f32x2 r2, f32x2 nd2,
                                            f32x2& nx, f32x2& ny, f32x2& ws) {
    float ax, bx, ay, by, az, bz;
    upk2(sub2(pk2(p.x, p.x), cx), ax, bx);                                            // :757  |p - c| per channel, both pixels of the pair
    upk2(sub2(pk2(p.y, p.y), cy), ay, by);
    upk2(sub2(pk2(p.z, p.z), cz), az, bz);
    const f32x2 dr = pk2(fmaxf(fmaxf(fabsf(ax), fabsf(ay)), fabsf(az)), fmaxf(fmaxf(fabsf(bx), fabsf(by)), fabsf(bz)));
    const f32x2 xx = mul2(dr, dr);
    float ta, tb;
    if (FAST_DIV) {
        const f32x2 q0 = mul2(xx, r2);
        const f32x2 q = fma2(fma2(q0, nd2, xx), r2, q0);
        upk2(mul2(q, pk2(1.4426950216293334961f, 1.4426950216293334961f)), ta, tb);
    } else {
        float xa, xb;
        upk2(xx, xa, xb);
        ta = __fmul_rn(__fdiv_rn(xa, a.neg_sig_r2), 1.4426950216293334961f);
        tb = __fmul_rn(__fdiv_rn(xb, a.neg_sig_r2), 1.4426950216293334961f);
    }
    // exp_ref on both halves (:758)
    const bool tinya = ta < -126.0f, tinyb = tb < -126.0f;
    if (tinya) ta = __fmul_rn(ta, 0.5f);
    if (tinyb) tb = __fmul_rn(tb, 0.5f);
    float ea = ex2_mufu(ta), eb = ex2_mufu(tb);
    if (tinya) ea = __fmul_rn(ea, ea);
    if (tinyb) eb = __fmul_rn(eb, eb);
    const f32x2 wgt = mul2(pk2(ea, eb), gg);                                          // :759-760
    nx = fma2(wgt, pk2(fl.x, fl.x), nx);                                              // :782-783
    ny = fma2(wgt, pk2(fl.y, fl.y), ny);
    ws = add2(ws, wgt);
}

template <bool FAST_DIV>
__global__ void __launch_bounds__(S4_TX* S4_TY) k_flow_smooth4(SmoothArgs a, const __grid_constant__ SmoothLut lut, const __grid_constant__ CUtensorMap tmap,
                                                               int use_tma) {
    extern __shared__ __align__(128) float4 smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    constexpr int R = S4_R, TW = S4_TW, TH = S4_TH, NT = S4_TX * S4_TY;
    float4* s_pix = smem;                                             // [TH][TW] colours of image 1
    float2* s_flow = reinterpret_cast<float2*>(smem + TW * TH);       // [TH][TW]
    float2* s_gg = s_flow + TW * TH;                                  // [pair][tile row t][|dx|]: spatial weights of the pair's two pixels
    const int b = blockIdx.z;
    const float2* f = a.src + (size_t)b * a.w * a.h;
    const float4* img = a.pix + (size_t)b * a.plane + (size_t)PAD * a.pw + PAD;
    const int x0 = blockIdx.x * S4_TX - R, y0 = a.y0 + blockIdx.y * S4_TY * S4_PY - R;
    const int tid = threadIdx.y * S4_TX + threadIdx.x;
    if (use_tma) {
        if (tid == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&s_bar, (unsigned)(TW * TH * sizeof(float4)));
            tma_load_3d(s_pix, &tmap, (x0 + PAD) * 4, y0 + PAD, b, &s_bar);
        }
    }
    // cBlfGaussian[|dx|] * cBlfGaussian[|dy|] (:759) for the two pixels k = 2*pair, 2*pair+1 of a thread at its tile row t: dy = t - R - k
    for (int i = tid; i < 2 * S4_ROWS * (R + 1); i += NT) {
        const int adx = i % (R + 1), t = (i / (R + 1)) % S4_ROWS, pr = i / ((R + 1) * S4_ROWS);
        const int dy0 = t - R - 2 * pr, dy1 = dy0 - 1;
        s_gg[i] = make_float2(abs(dy0) <= R ? __fmul_rn(lut.g[adx], lut.g[abs(dy0)]) : 0.f, abs(dy1) <= R ? __fmul_rn(lut.g[adx], lut.g[abs(dy1)]) : 0.f);
    }
    unsigned far_mask = 0;   // tile entries of this thread that must not contribute (outside the image / unknown flow)
    {
        int k = 0;
        for (int i = tid; i < TW * TH; i += NT, k++) {
            const int ty = i / TW, tx = i - ty * TW;
            const int cx = x0 + tx, cy = y0 + ty;
            float2 fl = make_float2(EPPM_UNKNOWN_FLOW, EPPM_UNKNOWN_FLOW);
            float4 p = make_float4(SMOOTH_FAR, SMOOTH_FAR, SMOOTH_FAR, 0.f);
            const bool inside = cx >= 0 && cy >= 0 && cx < a.w && cy < a.h;
            if (inside) fl = f[(size_t)cy * a.w + cx];
            const bool known = !(fmaxf(fl.x, fl.y) > EPPM_UNKNOWN_FLOW_THRESH);       // :778 (same truth table as x > T || y > T)
            if (!known) { fl = make_float2(0.f, 0.f); far_mask |= 1u << k; }
            else if (!use_tma) p = ldpix(img + (size_t)cy * a.pw + cx);
            s_flow[i] = fl;
            if (!use_tma) s_pix[i] = p;
        }
    }
    if (use_tma) {
        mbar_wait(&s_bar, 0);
        int k = 0;
        for (int i = tid; i < TW * TH; i += NT, k++)
            if (far_mask >> k & 1) s_pix[i] = make_float4(SMOOTH_FAR, SMOOTH_FAR, SMOOTH_FAR, 0.f);
    }
    __syncthreads();
    const int x = blockIdx.x * S4_TX + threadIdx.x;
    const int ly = threadIdx.y * S4_PY;
    const int y = a.y0 + blockIdx.y * S4_TY * S4_PY + ly;
    if (x >= a.w || y >= a.y1) return;
    // centre colours from the image itself (a centre with unknown flow is still filtered with its own colour, :764-771)
    f32x2 cx[2], cy[2], cz[2];
#pragma unroll
    for (int pr = 0; pr < 2; pr++) {
        const int ya = min(y + 2 * pr, a.h - 1), yb = min(y + 2 * pr + 1, a.h - 1);
        const float4 ca = ldpix(img + (size_t)ya * a.pw + x), cb = ldpix(img + (size_t)yb * a.pw + x);
        cx[pr] = pk2(ca.x, cb.x); cy[pr] = pk2(ca.y, cb.y); cz[pr] = pk2(ca.z, cb.z);
    }
    const f32x2 r2 = pk2(a.recip, a.recip), nd2 = pk2(-a.neg_sig_r2, -a.neg_sig_r2);
    f32x2 nx[2], ny[2], ws[2];
#pragma unroll
    for (int pr = 0; pr < 2; pr++) nx[pr] = ny[pr] = ws[pr] = pk2(0.f, 0.f);
    const int col = threadIdx.x + R;
    // tile rows 0,1: pair (A,B) only; rows 2 .. 2R+1: both pairs; rows 2R+2, 2R+3: pair (C,D) only
#pragma unroll 1
    for (int t = 0; t < S4_ROWS; t++) {
        const int rowb = (ly + t) * TW + col;
        const float2* g0 = s_gg + t * (R + 1);
        const float2* g1 = s_gg + (S4_ROWS + t) * (R + 1);
        if (t >= 2 && t <= 2 * R + 1) {
EPPM_PRAGMA(unroll S4_UNROLL)
            for (int dx = -R; dx <= R; dx++) {
                const float4 p = s_pix[rowb + dx];
                const float2 fl = s_flow[rowb + dx];
                const int adx = dx < 0 ? -dx : dx;
                const float2 ga = g0[adx], gb = g1[adx];
                smooth_tap2<FAST_DIV>(a, cx[0], cy[0], cz[0], p, fl, pk2(ga.x, ga.y), r2, nd2, nx[0], ny[0], ws[0]);
                smooth_tap2<FAST_DIV>(a, cx[1], cy[1], cz[1], p, fl, pk2(gb.x, gb.y), r2, nd2, nx[1], ny[1], ws[1]);
            }
        } else if (t < 2) {
EPPM_PRAGMA(unroll S4_UNROLL)
            for (int dx = -R; dx <= R; dx++) {
                const float2 ga = g0[dx < 0 ? -dx : dx];
                smooth_tap2<FAST_DIV>(a, cx[0], cy[0], cz[0], s_pix[rowb + dx], s_flow[rowb + dx], pk2(ga.x, ga.y), r2, nd2, nx[0], ny[0], ws[0]);
            }
        } else {
EPPM_PRAGMA(unroll S4_UNROLL)
            for (int dx = -R; dx <= R; dx++) {
                const float2 gb = g1[dx < 0 ? -dx : dx];
                smooth_tap2<FAST_DIV>(a, cx[1], cy[1], cz[1], s_pix[rowb + dx], s_flow[rowb + dx], pk2(gb.x, gb.y), r2, nd2, nx[1], ny[1], ws[1]);
            }
        }
    }
#pragma unroll
    for (int pr = 0; pr < 2; pr++) {
        float n0, n1, m0, m1, w0, w1;
        upk2(nx[pr], n0, n1); upk2(ny[pr], m0, m1); upk2(ws[pr], w0, w1);
        const int ya = y + 2 * pr;
        if (ya < a.y1) {
            float2 o = f[(size_t)ya * a.w + x];
            if (w0 != 0.f) o = make_float2(__fdiv_rn(n0, w0), __fdiv_rn(m0, w0));     // :790-796 (untouched otherwise)
            a.dst[(size_t)b * a.w * a.h + (size_t)ya * a.w + x] = o;
        }
        if (ya + 1 < a.y1) {
            float2 o = f[(size_t)(ya + 1) * a.w + x];
            if (w1 != 0.f) o = make_float2(__fdiv_rn(n1, w1), __fdiv_rn(m1, w1));
            a.dst[(size_t)b * a.w * a.h + (size_t)(ya + 1) * a.w + x] = o;
        }
    }
}

// Exhaustive check that the 3-instruction constant division equals div.rn for every float in [lo, hi) (bit patterns).
__global__ void k_selftest_const_div(float d, float r, unsigned lo_bits, unsigned hi_bits, unsigned long long* mismatches) {
    unsigned long long local = 0;
    for (unsigned long long b = lo_bits + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; b < hi_bits;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)b);
        const float q0 = __fmul_rn(x, r);
        const float q = __fmaf_rn(__fmaf_rn(q0, -d, x), r, q0);
        if (__float_as_uint(q) != __float_as_uint(__fdiv_rn(x, d))) local++;
    }
    if (local) atomicAdd(mismatches, local);
}

long long selftest_const_div(float d, unsigned lo_bits, unsigned hi_bits) {
    unsigned long long* dm = nullptr;
    if (cudaMalloc((void**)&dm, 8) != cudaSuccess) return -1;
    cudaMemset(dm, 0, 8);
    volatile float one = 1.0f;
    const float r = one / d;
    k_selftest_const_div<<<148 * 8, 256>>>(d, r, lo_bits, hi_bits, dm);
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpy(&h, dm, 8, cudaMemcpyDeviceToHost);
    cudaFree(dm);
    return e == cudaSuccess ? (long long)h : -1;
}

void op_refine(eppm_context* c, const float4* pix1, const float4* pix2, const LevelGeom& g, const float2* coarse, int ws, int hs, int upsample,
               float2* out, int n, int y0, int y1) {
    if (y1 < 0) y1 = g.h;
    RefineArgs a;
    a.pix1 = pix1;
    a.pix2 = pix2;
    a.plane = g.plane; a.pw = g.pw; a.w = g.w; a.h = g.h;
    a.coarse = coarse; a.ws = ws; a.hs = hs;
    a.flow = out;
    a.upsample = upsample;
    a.y0 = y0;
    a.px_bytes = (int)sizeof(float4);
    dim3 blk(RF_PIX * 3), grd((g.w + RF_PIX - 1) / RF_PIX, y1 - y0, n);
    if (!(c->variant & EPPM_VAR_REFINE_GENERIC)) {
        // table-driven kernel: the site table of this pitch and stride was built and verified at eppm_create
        const AffineTab* tabp = nullptr;
        for (int l = 0; l < c->n_levels; l++)
            if (c->aff_ok[l] && c->lv[l].pw == g.pw && c->lv[l].w >= g.w && c->lv[l].h >= g.h) tabp = &c->aff_tab[l];
        if (tabp) {
            const dim3 blk9(RF_PIX * 9);
            const int v = c->variant;
            switch (c->prm.patch_stride) {
            case 1: k_c2f_refine_tab<true, 3, RF_MINBLOCKS, 1><<<grd, blk, 0, c->stream>>>(a, c->cost_lut, *tabp); break;
            case 3: k_c2f_refine_tab<true, 3, RF_MINBLOCKS, 3><<<grd, blk, 0, c->stream>>>(a, c->cost_lut, *tabp); break;
            default:
                if (v & EPPM_VAR_REFINE_NOGROUP) k_c2f_refine_tab<false, 3, RF_MINBLOCKS, 2><<<grd, blk, 0, c->stream>>>(a, c->cost_lut, *tabp);
                else if (v & EPPM_VAR_REFINE_9WARP3) k_c2f_refine_tab<true, 1, 3, 2><<<grd, blk9, 0, c->stream>>>(a, c->cost_lut, *tabp);
                else if (v & EPPM_VAR_REFINE_PK_BRANCH) k_c2f_refine_pk<RF_PK_MINBLOCKS, 2, false><<<grd, blk, 0, c->stream>>>(a, c->cost_lut, *tabp);
                else if (v & EPPM_VAR_REFINE_PK) k_c2f_refine_pk<RF_PK_MINBLOCKS, 2, true><<<grd, blk, 0, c->stream>>>(a, c->cost_lut, *tabp);
                else {
                    // Default (mode 23): warp = candidate row, census table at the shared-window base, the CTA's image-1 samples staged in shared memory by ONE
                    // TMA copy (50 pixels x every second of 19 rows; 7.75 ms per 1080p pair at level 0 against 7.85 for mode 10, which reads them through L1), and the
                    // __expf fix-up test dropped behind an exact first patch row where that provably changes no bit (FASTW: 7.75 -> 7.54).  Measured per 1080p pair at level 0 (round 2,
                    // tools/variant_times.py, 16 pairs): column kernel 8.28 ms; its knobs allrows 9.72, wide address 8.78, 6 CTAs 8.54; row kernel
                    // 8.06-8.12, + fixed-address census table 7.85, + guard-free loop for interior warps 7.92 (spills), warp-uniform guards (mode 16) 7.88,
                    // census table indexed by the XOR byte instead of POPC (modes 12-15: plain 7.94, replicated x8 / x16 / x32 8.23 / 8.27 / 12.3).
                    // Tuning knob EPPM_REFINE_MODE: 0-7 = column kernel with allrows + 2 * wide + 4 * (6 CTAs per SM), 8-11 = row kernel + 2 * table at base + guard-free
                    static const int mode = getenv("EPPM_REFINE_MODE") ? atoi(getenv("EPPM_REFINE_MODE")) : 23;
                    static const CUtensorMap dummy_map = {};
                    const int md = (v & EPPM_VAR_REFINE_COLUMN) ? 0 : (v & EPPM_VAR_REFINE_VOLUME) ? 20 : ((mode == 19 || mode == 23) && (v & EPPM_VAR_REFINE_NOFASTW)) ? 18 : mode;
#define EPPM_RT(MB, AR, WD) k_c2f_refine_tab<true, 3, MB, 2, AR, WD><<<grd, blk, 0, c->stream>>>(a, c->cost_lut, *tabp)
                    switch (md) {
#define EPPM_RR(L0, FP) k_c2f_refine_row<7, 2, L0, FP><<<grd, blk, (16 + 9 * RF_PIX) * sizeof(float), c->stream>>>(a, c->cost_lut, *tabp, dummy_map)
                    case 8: EPPM_RR(false, false); break;
                    case 9: EPPM_RR(false, true); break;
                    case 20: case 21: case 22: {   // shared AD + census volume (6 / 7 / 5 CTAs per SM); needs everything mode 19 needs
                        int lvl = -1;
                        for (int l = 0; l < c->n_levels; l++)
                            if (pix1 == c->pix[0][l] && c->tmap_refine_ok[l]) lvl = l;
                        if (lvl >= 0 && c->vol_ok && lut0_window_base_ok(c->device)) {
                            static bool attr_v[64] = {};
                            if (!attr_v[c->device & 63]) {
                                cudaFuncSetAttribute(k_c2f_refine_vol<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, RF_VOL_SMEM);
                                cudaFuncSetAttribute(k_c2f_refine_vol<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, RF_VOL_SMEM);
                                cudaFuncSetAttribute(k_c2f_refine_vol<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, RF_VOL_SMEM);
                                attr_v[c->device & 63] = true;
                            }
                            if (md == 21) k_c2f_refine_vol<7><<<grd, blk, RF_VOL_SMEM, c->stream>>>(a, c->cost_lut, *tabp, c->tmap_refine[lvl], c->vol_tab);
                            else if (md == 22) k_c2f_refine_vol<5><<<grd, blk, RF_VOL_SMEM, c->stream>>>(a, c->cost_lut, *tabp, c->tmap_refine[lvl], c->vol_tab);
                            else k_c2f_refine_vol<6><<<grd, blk, RF_VOL_SMEM, c->stream>>>(a, c->cost_lut, *tabp, c->tmap_refine[lvl], c->vol_tab);
                            break;
                        }
                    }
                    // fall through
                    case 24: case 25:   // mode 23 compiled for 8 / 6 CTAs per SM
                    case 23:     // default: mode 19 + no validity guards in the fix-up-free loop of warps whose 96 candidates are all valid (7.54 -> 7.37 ms)
                    case 19:     // mode 18 + fix-up-free loop behind an exact first patch row (FASTW)
                    case 18: {   // image-1 tile staged by TMA (needs the level's refine tensor map and the table-at-base addressing); else mode 10
                        int lvl = -1;
                        for (int l = 0; l < c->n_levels; l++)
                            if (pix1 == c->pix[0][l] && c->tmap_refine_ok[l]) lvl = l;
                        if (lvl >= 0 && lut0_window_base_ok(c->device)) {
                            if (md == 24) k_c2f_refine_row<8, 2, true, true, 0, 0, true, true><<<grd, blk, 1280 + RF_TILE_W * RF_TILE_H * sizeof(float4) + 16, c->stream>>>(a, c->cost_lut, *tabp, c->tmap_refine[lvl]);
                            else if (md == 25) k_c2f_refine_row<6, 2, true, true, 0, 0, true, true><<<grd, blk, 1280 + RF_TILE_W * RF_TILE_H * sizeof(float4) + 16, c->stream>>>(a, c->cost_lut, *tabp, c->tmap_refine[lvl]);
                            else if (md == 23) k_c2f_refine_row<7, 2, true, true, 0, 0, true, true><<<grd, blk, 1280 + RF_TILE_W * RF_TILE_H * sizeof(float4) + 16, c->stream>>>(a, c->cost_lut, *tabp, c->tmap_refine[lvl]);
                            else if (md == 19) k_c2f_refine_row<7, 2, true, false, 0, 0, true, true><<<grd, blk, 1280 + RF_TILE_W * RF_TILE_H * sizeof(float4) + 16, c->stream>>>(a, c->cost_lut, *tabp, c->tmap_refine[lvl]);
                            else k_c2f_refine_row<7, 2, true, false, 0, 0, true><<<grd, blk, 1280 + RF_TILE_W * RF_TILE_H * sizeof(float4) + 16, c->stream>>>(a, c->cost_lut, *tabp, c->tmap_refine[lvl]);
                            break;
                        }
                    }
                    // fall through
                    case 10: if (lut0_window_base_ok(c->device)) EPPM_RR(true, false); else EPPM_RR(false, false); break;
                    case 11: if (lut0_window_base_ok(c->device)) EPPM_RR(true, true); else EPPM_RR(false, true); break;
                    case 16: if (lut0_window_base_ok(c->device)) k_c2f_refine_row<7, 2, true, false, 0, 1><<<grd, blk, (16 + 9 * RF_PIX) * sizeof(float), c->stream>>>(a, c->cost_lut, *tabp, dummy_map); else EPPM_RR(false, false); break;
#define EPPM_RX(REP) { static bool at##REP[64] = {}; const size_t sm = (256 * REP + 9 * RF_PIX) * sizeof(float); \
                       if (!at##REP[c->device & 63]) { cudaFuncSetAttribute(k_c2f_refine_row<7, 2, true, false, REP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); at##REP[c->device & 63] = true; } \
                       k_c2f_refine_row<7, 2, true, false, REP><<<grd, blk, sm, c->stream>>>(a, c->cost_lut, *tabp, dummy_map); }
                    case 12: if (lut0_window_base_ok(c->device)) EPPM_RX(1) else EPPM_RR(false, false); break;
                    case 13: if (lut0_window_base_ok(c->device)) EPPM_RX(8) else EPPM_RR(false, false); break;
                    case 14: if (lut0_window_base_ok(c->device)) EPPM_RX(16) else EPPM_RR(false, false); break;
                    case 15: if (lut0_window_base_ok(c->device)) EPPM_RX(32) else EPPM_RR(false, false); break;
#undef EPPM_RX
#undef EPPM_RR
                    case 1: EPPM_RT(RF_TAB2_MINBLOCKS, true, false); break;
                    case 2: EPPM_RT(RF_TAB2_MINBLOCKS, false, true); break;
                    case 3: EPPM_RT(RF_TAB2_MINBLOCKS, true, true); break;
                    case 4: EPPM_RT(6, false, false); break;
                    case 5: EPPM_RT(6, true, false); break;
                    case 6: EPPM_RT(6, false, true); break;
                    case 7: EPPM_RT(6, true, true); break;
                    default: EPPM_RT(RF_TAB2_MINBLOCKS, false, false); break;
                    }
#undef EPPM_RT
                }
            }
            EPPM_LAUNCH_COUNT(1);
            return;
        }
    }
    switch (c->prm.patch_stride) {  // sample stride is a compile-time constant of the kernel
    case 1: k_c2f_refine<1><<<grd, blk, 0, c->stream>>>(a, c->cost_lut); break;
    case 3: k_c2f_refine<3><<<grd, blk, 0, c->stream>>>(a, c->cost_lut); break;
    default: k_c2f_refine<2><<<grd, blk, 0, c->stream>>>(a, c->cost_lut); break;
    }
    EPPM_LAUNCH_COUNT(1);
}

void op_smooth(eppm_context* c, const float2* src, float2* dst, const float4* pix1, const LevelGeom& g, int n, int y0, int y1) {
    if (y1 < 0) y1 = g.h;
    SmoothArgs a;
    a.y0 = y0; a.y1 = y1;
    a.src = src; a.dst = dst;
    a.pix = pix1;
    a.plane = g.plane; a.pw = g.pw; a.w = g.w; a.h = g.h;
    a.R = 2 * c->prm.blf_sig_s;
    a.neg_sig_r2 = -(c->prm.blf_sig_r * c->prm.blf_sig_r);
    volatile float one = 1.0f;
    a.recip = one / a.neg_sig_r2;
    a.fast_div = c->smooth_fast_div;
    // TMA path: the plane must be one of the context's own image-1 planes (a tensor map exists per level, built for one tile shape)
    int level = -1;
    for (int l = 0; l < c->n_levels; l++)
        if (pix1 == c->pix[0][l]) level = l;
    static const CUtensorMap dummy = {};
    if (a.R == S4_R && c->prm.blf_sig_r <= 10.f && !(c->variant & EPPM_VAR_SMOOTH_2ROW)) {
        // four rows per thread, packed pairs (SMOOTH_FAR = 1e4 needs exp(-(1e4/sig_r)^2) == 0, true for any sig_r <= 10)
        // the opt-in to > 48 KB of dynamic shared memory is per device: once per device the library has seen (a process may hold contexts on several)
        static bool attr4[64] = {};
        if (c->device < 0 || c->device >= 64 || !attr4[c->device]) {
            cudaFuncSetAttribute(k_flow_smooth4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S4_SMEM);
            cudaFuncSetAttribute(k_flow_smooth4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S4_SMEM);
            if (c->device >= 0 && c->device < 64) attr4[c->device] = true;
        }
        const int use_tma = level >= 0 && c->tmap_ok[level] && c->tmap_box_h == S4_TH;
        dim3 blk(S4_TX, S4_TY), grd((g.w + S4_TX - 1) / S4_TX, (y1 - y0 + S4_TY * S4_PY - 1) / (S4_TY * S4_PY), n);
        if (a.fast_div) k_flow_smooth4<true><<<grd, blk, S4_SMEM, c->stream>>>(a, c->smooth_lut, use_tma ? c->tmap_pix0[level] : dummy, use_tma);
        else k_flow_smooth4<false><<<grd, blk, S4_SMEM, c->stream>>>(a, c->smooth_lut, use_tma ? c->tmap_pix0[level] : dummy, use_tma);
        EPPM_LAUNCH_COUNT(1);
        return;
    }
    const int TW = SM_TX + 2 * a.R, TH = SM_TY * SM_PY + 2 * a.R;
    const size_t smem = (size_t)TW * TH * (sizeof(float4) + sizeof(float2)) + (size_t)(a.R + 1) * (a.R + 1) * sizeof(float);
    static bool attr_set[64] = {};
    if (c->device < 0 || c->device >= 64 || !attr_set[c->device]) {
        cudaFuncSetAttribute(k_flow_smooth, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (c->device >= 0 && c->device < 64) attr_set[c->device] = true;
    }
    dim3 blk(SM_TX, SM_TY), grd((g.w + SM_TX - 1) / SM_TX, (y1 - y0 + SM_TY * SM_PY - 1) / (SM_TY * SM_PY), n);
    const int use_tma = level >= 0 && c->tmap_ok[level] && c->tmap_box_h == TH && TW * 4 <= 256 && TH <= 256;
    k_flow_smooth<<<grd, blk, smem, c->stream>>>(a, c->smooth_lut, use_tma ? c->tmap_pix0[level] : dummy, use_tma);
    EPPM_LAUNCH_COUNT(1);
}

// d_bilateral_upsample_flow (bao_pmflow_refine_kernel.cu:829-865), declared but never called by the reference's pipeline (its call in
// baoCudaBLF_C2F is commented out, :1080): joint-bilateral UPSAMPLING of a coarser flow -- the taps of the smoothing filter, but each tap
// reads the small flow at (int(cy / ratio), int(cx / ratio)) and the result is scaled by the ratio.  Out of place, hence deterministic.
__global__ void __launch_bounds__(256) k_flow_bilateral_upsample(SmoothArgs a, const __grid_constant__ SmoothLut lut, const float2* __restrict__ small, int ws,
                                                                 float ratio) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.w || y >= a.h) return;
    const float4* img = a.pix + (size_t)PAD * a.pw + PAD;
    const float4 c = ldpix(img + (size_t)y * a.pw + x);
    const float r = a.recip, nd = -a.neg_sig_r2;
    float nx = 0.f, ny = 0.f, wsum = 0.f;
    for (int dy = -a.R; dy <= a.R; dy++) {
        const int cy = y + dy;
        if (cy < 0 || cy >= a.h) continue;
        const int sy = __float2int_rz(__fdiv_rn((float)cy, ratio));
        const float gy = lut.g[abs(dy)];
        for (int dx = -a.R; dx <= a.R; dx++) {
            const int cx = x + dx;
            if (cx < 0 || cx >= a.w) continue;
            const float2 fl = small[(size_t)sy * ws + __float2int_rz(__fdiv_rn((float)cx, ratio))];
            if (fl.x > EPPM_UNKNOWN_FLOW_THRESH || fl.y > EPPM_UNKNOWN_FLOW_THRESH) continue;
            smooth_tap(a, c, ldpix(img + (size_t)cy * a.pw + cx), fl, __fmul_rn(lut.g[abs(dx)], gy), r, nd, nx, ny, wsum);
        }
    }
    if (wsum != 0.f)
        a.dst[(size_t)y * a.w + x] = make_float2(__fmul_rn(__fdiv_rn(nx, wsum), ratio), __fmul_rn(__fdiv_rn(ny, wsum), ratio));
}

// d_image_bilateral_filtering (bao_pmflow_refine_kernel.cu:976-1020): the guide image filtered by its own joint-bilateral weights.
// baoCudaImageSmoothing (:1022-1057) runs a 5x5 median into the same output first; the bilateral pass reads the ORIGINAL image and
// overwrites every pixel, so the median is dead work and is not executed here.  Alpha is left uninitialised by the reference; 0 here.
__global__ void __launch_bounds__(256) k_image_bilateral(SmoothArgs a, const __grid_constant__ SmoothLut lut, const uchar4* __restrict__ src,
                                                         uchar4* __restrict__ dst, size_t img_w) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.w || y >= a.h) return;
    const float4* img = a.pix + (size_t)PAD * a.pw + PAD;
    const float4 c = ldpix(img + (size_t)y * a.pw + x);
    const float r = a.recip, nd = -a.neg_sig_r2;
    float nr = 0.f, ng = 0.f, nb = 0.f, wsum = 0.f;
    for (int dy = -a.R; dy <= a.R; dy++) {
        const int cy = y + dy;
        if (cy < 0 || cy >= a.h) continue;
        const float gy = lut.g[abs(dy)];
        for (int dx = -a.R; dx <= a.R; dx++) {
            const int cx = x + dx;
            if (cx < 0 || cx >= a.w) continue;
            const float dr = max3abs_diff(ldpix(img + (size_t)cy * a.pw + cx), c);    // :757
            const float xx = __fmul_rn(dr, dr);
            float q;
            if (a.fast_div) {
                const float q0 = __fmul_rn(xx, r);
                q = __fmaf_rn(__fmaf_rn(q0, nd, xx), r, q0);
            } else {
                q = __fdiv_rn(xx, a.neg_sig_r2);
            }
            const float wgt = __fmul_rn(exp_ref(q), __fmul_rn(lut.g[abs(dx)], gy));   // :758-760
            const uchar4 p = src[(size_t)cy * img_w + cx];
            nr = __fmaf_rn(wgt, (float)p.x, nr);                                         // :995-997
            ng = __fmaf_rn(wgt, (float)p.y, ng);
            nb = __fmaf_rn(wgt, (float)p.z, nb);
            wsum = __fadd_rn(wsum, wgt);
        }
    }
    uchar4 o = src[(size_t)y * img_w + x];
    if (wsum != 0.f)
        o = make_uchar4((unsigned char)__float2uint_rz(__fdiv_rn(nr, wsum)), (unsigned char)__float2uint_rz(__fdiv_rn(ng, wsum)),
                        (unsigned char)__float2uint_rz(__fdiv_rn(nb, wsum)), 0);
    dst[(size_t)y * img_w + x] = o;
}

void op_image_bilateral(eppm_context* c, uchar4* dst, const uchar4* src, size_t pitch_bytes, const float4* pix1, const LevelGeom& g) {
    SmoothArgs a = {};
    a.pix = pix1;
    a.plane = g.plane; a.pw = g.pw; a.w = g.w; a.h = g.h;
    a.R = 2 * c->prm.blf_sig_s;
    a.neg_sig_r2 = -(c->prm.blf_sig_r * c->prm.blf_sig_r);
    volatile float one = 1.0f;
    a.recip = one / a.neg_sig_r2;
    a.fast_div = c->smooth_fast_div;
    k_image_bilateral<<<dim3((g.w + 31) / 32, (g.h + 7) / 8), dim3(32, 8), 0, c->stream>>>(a, c->smooth_lut, src, dst, pitch_bytes / sizeof(uchar4));
    EPPM_LAUNCH_COUNT(1);
}

void op_flow_bilateral_upsample(eppm_context* c, float2* dst, const float4* pix1, const LevelGeom& g, const float2* small, int ws, float ratio) {
    SmoothArgs a = {};
    a.dst = dst; a.pix = pix1;
    a.plane = g.plane; a.pw = g.pw; a.w = g.w; a.h = g.h;
    a.R = 2 * c->prm.blf_sig_s;
    a.neg_sig_r2 = -(c->prm.blf_sig_r * c->prm.blf_sig_r);
    volatile float one = 1.0f;
    a.recip = one / a.neg_sig_r2;
    a.fast_div = c->smooth_fast_div;
    a.y0 = 0; a.y1 = g.h;
    k_flow_bilateral_upsample<<<dim3((g.w + 31) / 32, (g.h + 7) / 8), dim3(32, 8), 0, c->stream>>>(a, c->smooth_lut, small, ws, ratio);
    EPPM_LAUNCH_COUNT(1);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link dependency on libcuda).
// Map of the image-1 packed planes of one level: dims (x in floats = pw*4, y = ph, z = plane index), box = smoothing tile.
bool build_smooth_tensor_maps(eppm_context* c) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    // tile of the kernel op_smooth will pick: four rows per thread when R = 10, else the generic two-row kernel
    const int R = 2 * c->prm.blf_sig_s;
    const bool four = R == S4_R && c->prm.blf_sig_r <= 10.f && !(c->variant & EPPM_VAR_SMOOTH_2ROW);
    const int TW = four ? S4_TW : SM_TX + 2 * R, TH = four ? S4_TH : SM_TY * SM_PY + 2 * R;
    c->tmap_box_h = TH;
    for (int l = 0; l < c->n_levels; l++) {
        c->tmap_ok[l] = 0;
        if (TW * 4 > 256 || TH > 256) continue;
        const LevelGeom& g = c->lv[l];
        const cuuint64_t dims[3] = {(cuuint64_t)g.pw * 4, (cuuint64_t)g.ph, (cuuint64_t)c->max_batch};
        const cuuint64_t strides[2] = {(cuuint64_t)g.pw * 16, (cuuint64_t)g.plane * 16};  // bytes, dims 1 and 2
        const cuuint32_t box[3] = {(cuuint32_t)TW * 4, (cuuint32_t)TH, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = ((encode_fn)fn)(&c->tmap_pix0[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, c->pix[0][l], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        c->tmap_ok[l] = r == CUDA_SUCCESS;
        // image-1 tile of the refine kernel (EPPM_REFINE_MODE=18): 50 pixels x 19 rows visited with an element stride of 2 in y = 10 rows
        const cuuint32_t rbox[3] = {(cuuint32_t)RF_TILE_W * 4, (cuuint32_t)(2 * PATCH_R + 1), 1};
        const cuuint32_t restr[3] = {1, 2, 1};
        r = ((encode_fn)fn)(&c->tmap_refine[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, c->pix[0][l], dims, strides, rbox, restr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        c->tmap_refine_ok[l] = r == CUDA_SUCCESS;
        // census + pack: 32 x 8 output pixels + a one-pixel halo of the dense RGBA level (uchar4 as one 32-bit element); the box is 40 wide because it
        // must start on a 16-byte boundary (see k_census_pack_tile).  Needs 16-byte row and plane strides.
        for (int img = 0; img < 2; img++) {
            c->tmap_rgba_ok[img][l] = 0;
            if (g.w % 4 || ((size_t)g.w * g.h) % 4 || !c->rgba[img][l]) continue;
            const cuuint64_t cdims[3] = {(cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)c->max_batch};
            const cuuint64_t cstrides[2] = {(cuuint64_t)g.w * 4, (cuuint64_t)g.w * g.h * 4};
            const cuuint32_t cbox[3] = {CEN_TILE_W, CEN_TILE_H, 1};
            r = ((encode_fn)fn)(&c->tmap_rgba[img][l], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, c->rgba[img][l], cdims, cstrides, cbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            c->tmap_rgba_ok[img][l] = r == CUDA_SUCCESS;
        }
    }
    return true;
}

// rows of level `level` that correspond to the context's band of coarsest-level rows (the last band takes the remainder)
void band_rows(const eppm_context* c, int level, int* y0, int* y1) {
    const int L = c->n_levels - 1, sh = L - level;
    *y0 = c->band_y0 << sh;
    *y1 = c->band_y1 >= c->lv[L].h ? c->lv[level].h : min(c->lv[level].h, c->band_y1 << sh);
}

// One step of the coarse-to-fine stage on the context's band: kind 0 = x2 upsample + plane-fitting refine of `level`
// (flow[level+1] -> flow_tmp), 1 = smoothing flow_tmp -> flow[level], 2 = final smoothing flow[0] -> out (or flow_tmp).
void run_c2f_step(eppm_context* c, int level, int kind, float2* out) {
    const int n = c->n_cur;
    int y0, y1;
    band_rows(c, level, &y0, &y1);
    const LevelGeom& g = c->lv[level];
    if (kind == 0) {
        const LevelGeom& gs = c->lv[level + 1];
        if (c->profile && level == 0) cudaEventRecord(c->ev_k[0], c->stream);
        op_refine(c, c->pix[0][level], c->pix[1][level], g, c->flow[level + 1], gs.w, gs.h, 1, c->flow_tmp, n, y0, y1);
        if (c->profile && level == 0) cudaEventRecord(c->ev_k[1], c->stream);
    } else if (c->inplace) {
        // reference order: the smoothing rewrites the plane it reads (whole level; not combined with spatial tiling)
        const size_t bytes = (size_t)n * g.w * g.h * sizeof(float2);
        float2* dst = kind == 1 ? c->flow[level] : (out ? out : c->flow_tmp);
        const float2* src = kind == 1 ? c->flow_tmp : c->flow[0];
        cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream);
        op_smooth_inplace(c, dst, c->pix[0][level], g, n);
    } else if (kind == 1) {
        op_smooth(c, c->flow_tmp, c->flow[level], c->pix[0][level], g, n, y0, y1);
    } else {
        if (c->profile) cudaEventRecord(c->ev_k[2], c->stream);
        op_smooth(c, c->flow[0], out ? out : c->flow_tmp, c->pix[0][0], g, n, y0, y1);
        if (c->profile) cudaEventRecord(c->ev_k[3], c->stream);
    }
}

void run_c2f(eppm_context* c, float* d_flow_out) {
    for (int level = c->n_levels - 2; level >= 0; level--) {
        run_c2f_step(c, level, 0, nullptr);
        if (level == 0 && c->prm.subpixel_final) op_subpix_final(c);   // opt-in: sub-pixel offsets on the integer flow of the level-0 refine
        run_c2f_step(c, level, 1, nullptr);
    }
    // final smoothing at level 0 (…cuda.cpp:289); with a single level the loop above did not run
    float2* out = reinterpret_cast<float2*>(d_flow_out);
    run_c2f_step(c, 0, 2, out);
    if (!out) cudaMemcpyAsync(c->flow[0], c->flow_tmp, (size_t)c->n_cur * c->lv[0].w * c->lv[0].h * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream);
}

}  // namespace eppm
