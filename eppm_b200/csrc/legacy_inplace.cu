// Opt-in "in-place" forms of the three neighbourhood filters the reference runs IN PLACE: outlier removal
// (bao_pmflow_refine_kernel.cu:149-193), the weighted median (:206-286) and the joint-bilateral flow smoothing (:764-826).
//
// The reference reads its neighbours from the array it is writing, so what a thread sees depends on which warps have already
// finished: its output is a function of the launch geometry and of warp scheduling.  The default kernels of this library give those
// passes snapshot semantics (consistency.cu, refine.cu).  The kernels here instead keep the reference's structure -- 16 x 16
// blocks with the reference's block -> pixel map, one thread per pixel walking its window in the reference's raster order,
// neighbours re-read from the array being updated, the pixel's own result stored at the end -- so that on the same GPU the same kind
// of warps finish first.  Per-tap arithmetic is the bit-exact restatement used by the default kernels.  Selected by
// eppm_params::inplace_filters (or EPPM_INPLACE_LEGACY=1 in the environment); DESIGN.md §3 reports how far this closes the gap
// to the reference build and why it cannot close it completely.
#include <float.h>

#include "eppm_internal.h"

namespace eppm {

constexpr int LB = 16;   // BLOCK_DIM_X / BLOCK_DIM_Y of the reference (bao_pmflow_refine_kernel.cu:42-43)

// d_outlier_removal (:149-182)
__global__ void __launch_bounds__(LB* LB) k_outlier_inplace(short2* nnf, float* cost, int w, int h, int R, int sim, int count_thresh) {
    const int x = blockIdx.x * LB + threadIdx.x, y = blockIdx.y * LB + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t off = (size_t)blockIdx.z * w * h;
    volatile short2* f = nnf + off;   // volatile: every neighbour is read from memory when the loop reaches it
    short2 cur;
    cur.x = f[y * w + x].x; cur.y = f[y * w + x].y;
    if (cur.x < 0 && cur.y < 0) return;   // :156 skip occlusion
    cur.x -= x; cur.y -= y;
    int count = 0;
    for (int dy = -R; dy <= R; dy++)
        for (int dx = -R; dx <= R; dx++) {
            const int cx = x + dx, cy = y + dy;
            if (cx < 0 || cy < 0 || cx >= w || cy >= h) continue;
            short2 nb;
            nb.x = f[cy * w + cx].x; nb.y = f[cy * w + cx].y;
            nb.x -= cx; nb.y -= cy;
            if (abs(nb.x - cur.x) <= sim && abs(nb.y - cur.y) <= sim) count++;
        }
    if (count < count_thresh) {
        nnf[off + (size_t)y * w + x] = make_short2(INVALID_LOCATION, INVALID_LOCATION);
        cost[off + (size_t)y * w + x] = FLT_MAX;
    }
}

// d_weighted_median_filtering (:206-259): for every valid candidate of the window the whole window is read again
__global__ void __launch_bounds__(LB* LB) k_wmf_inplace(short2* nnf, const float4* __restrict__ pix, size_t plane, int pw, int w, int h, int R,
                                                        bool only_occlusion, float neg_sig_r2, const __grid_constant__ WmfLut lut) {
    const int x = blockIdx.x * LB + threadIdx.x, y = blockIdx.y * LB + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t off = (size_t)blockIdx.z * w * h;
    volatile short2* f = nnf + off;
    short2 out;
    out.x = f[y * w + x].x; out.y = f[y * w + x].y;
    if (only_occlusion && out.x >= 0 && out.y >= 0) return;
    const float4* img = pix + (size_t)blockIdx.z * plane + (size_t)PAD * pw + PAD;
    const float4 c = ldpix(img + (size_t)y * pw + x);
    float best = FLT_MAX;
    for (int dy = -R; dy <= R; dy++)
        for (int dx = -R; dx <= R; dx++) {
            const int cy = y + dy, cx = x + dx;
            if (cx < 0 || cy < 0 || cx >= w || cy >= h) continue;
            short2 cand;
            cand.x = f[cy * w + cx].x; cand.y = f[cy * w + cx].y;
            if (cand.x < 0 || cand.y < 0) continue;
            cand.x -= cx; cand.y -= cy;
            float cost_sum = 0.f, weight_sum = 0.f;
            for (int dy2 = -R; dy2 <= R; dy2++)
                for (int dx2 = -R; dx2 <= R; dx2++) {
                    const int cy2 = y + dy2, cx2 = x + dx2;
                    if (cx2 < 0 || cy2 < 0 || cx2 >= w || cy2 >= h) continue;
                    short2 cur;
                    cur.x = f[cy2 * w + cx2].x; cur.y = f[cy2 * w + cx2].y;
                    if (cur.x < 0 || cur.y < 0) continue;
                    cur.x -= cx2; cur.y -= cy2;
                    const float dr = max3abs_diff(ldpix(img + (size_t)cy2 * pw + cx2), c);
                    const float coef_r = __expf(__fdiv_rn(__fmul_rn(dr, dr), neg_sig_r2));   // :198-204
                    const float wk = __fmul_rn(coef_r, __fmul_rn(lut.g[abs(dx2)], lut.g[abs(dy2)]));
                    const int dist = max(abs(cand.x - cur.x), abs(cand.y - cur.y));
                    cost_sum = __fmaf_rn(wk, (float)dist, cost_sum);   // :244
                    weight_sum = __fadd_rn(weight_sum, wk);
                }
            if (weight_sum > 0.0f && cost_sum < best) {
                best = cost_sum;
                out.x = cand.x + x; out.y = cand.y + y;
            }
        }
    if (out.x < 0 || out.y < 0) return;   // :257
    nnf[off + (size_t)y * w + x] = out;
}

// d_flow_bilateral_filtering (:764-799)
__global__ void __launch_bounds__(LB* LB) k_smooth_inplace(float2* flow, const float4* __restrict__ pix, size_t plane, int pw, int w, int h, int R,
                                                           float neg_sig_r2, float recip, int fast_div, const __grid_constant__ SmoothLut lut) {
    const int x = blockIdx.x * LB + threadIdx.x, y = blockIdx.y * LB + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t off = (size_t)blockIdx.z * w * h;
    volatile float2* f = flow + off;
    const float4* img = pix + (size_t)blockIdx.z * plane + (size_t)PAD * pw + PAD;
    const float4 c = ldpix(img + (size_t)y * pw + x);
    float nx = 0.f, ny = 0.f, ws = 0.f;
    const float nd = -neg_sig_r2;
    for (int dy = -R; dy <= R; dy++)
        for (int dx = -R; dx <= R; dx++) {
            const int cy = y + dy, cx = x + dx;
            if (cx < 0 || cy < 0 || cx >= w || cy >= h) continue;
            float2 fl;
            fl.x = f[cy * w + cx].x; fl.y = f[cy * w + cx].y;
            if (fl.x > EPPM_UNKNOWN_FLOW_THRESH || fl.y > EPPM_UNKNOWN_FLOW_THRESH) continue;
            const float dr = max3abs_diff(ldpix(img + (size_t)cy * pw + cx), c);   // :757
            const float xx = __fmul_rn(dr, dr);
            float q;
            if (fast_div) {
                const float q0 = __fmul_rn(xx, recip);
                q = __fmaf_rn(__fmaf_rn(q0, nd, xx), recip, q0);
            } else {
                q = __fdiv_rn(xx, neg_sig_r2);
            }
            const float wgt = __fmul_rn(exp_ref(q), __fmul_rn(lut.g[abs(dx)], lut.g[abs(dy)]));   // :758-760
            nx = __fmaf_rn(wgt, fl.x, nx);   // :782-783
            ny = __fmaf_rn(wgt, fl.y, ny);
            ws = __fadd_rn(ws, wgt);
        }
    if (ws != 0.f) flow[off + (size_t)y * w + x] = make_float2(__fdiv_rn(nx, ws), __fdiv_rn(ny, ws));   // :790-796
}

void op_outlier_inplace(eppm_context* c, short2* nnf, float* cost, int w, int h, int n) {
    const int R = c->prm.stat_radius;
    dim3 blk(LB, LB), grd((w + LB - 1) / LB, (h + LB - 1) / LB, n);
    k_outlier_inplace<<<grd, blk, 0, c->stream>>>(nnf, cost, w, h, R, c->prm.stat_sim_thresh, (2 * R + 1) * (2 * R + 1) / 2);
    EPPM_LAUNCH_COUNT(1);
}

void op_wmf_inplace(eppm_context* c, short2* nnf, const float4* pix, size_t plane, int pw, int w, int h, int n, int iters, bool only_occlusion) {
    const float sr = c->prm.wmf_sig_r;
    dim3 blk(LB, LB), grd((w + LB - 1) / LB, (h + LB - 1) / LB, n);
    for (int it = 0; it < iters; it++) k_wmf_inplace<<<grd, blk, 0, c->stream>>>(nnf, pix, plane, pw, w, h, c->prm.wmf_radius, only_occlusion, -(sr * sr), c->wmf_lut);
    EPPM_LAUNCH_COUNT(iters);
}

void op_smooth_inplace(eppm_context* c, float2* flow, const float4* pix1, const LevelGeom& g, int n) {
    const float nsr2 = -(c->prm.blf_sig_r * c->prm.blf_sig_r);
    volatile float one = 1.0f;
    dim3 blk(LB, LB), grd((g.w + LB - 1) / LB, (g.h + LB - 1) / LB, n);
    k_smooth_inplace<<<grd, blk, 0, c->stream>>>(flow, pix1, g.plane, g.pw, g.w, g.h, 2 * c->prm.blf_sig_s, nsr2, one / nsr2, c->smooth_fast_div, c->smooth_lut);
    EPPM_LAUNCH_COUNT(1);
}

}  // namespace eppm
