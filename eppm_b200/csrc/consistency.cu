// Stage "consistency": left/right check, outlier removal, weighted-median filling of occlusions, scan-line hole
// filling and NNF -> flow at the coarsest level.  Restates bao_flow_patchmatch_multiscale_cuda.cpp:233-258 and
// bao_pmflow_refine_kernel.cu:53-92 (LR check), :146-193 (outlier removal), :198-286 (weighted median),
// :291-390 (hole filling), :636-655,724-734 (NNF -> flow).
//
// Semantics: the reference's neighbourhood filters update the field IN PLACE while other threads still read it
// (outlier removal :171-181, weighted median :224-258, hole filling :312-370), so its own output depends on warp
// scheduling.  Here each of those passes reads a snapshot and writes a second buffer (Jacobi semantics), which is the
// behaviour the reference converges to when every thread reads before any thread writes; DESIGN.md quantifies the
// difference against the reference build.
//
// B200 design: the weighted median is the only heavy part (up to 81 candidates x 81 taps per occluded pixel, 20
// sweeps).  Occluded pixels are compacted into a list once per sweep and each one is handled by a full warp: the
// 81 bilateral weights are computed once per pixel into shared memory and reused by all candidates, lanes take
// candidates round-robin and accumulate their own cost in the reference's tap order, and a shuffle arg-min picks the
// first minimum in candidate order (strict '<' of the reference).
#include <float.h>

#include "eppm_internal.h"

namespace eppm {

struct FieldArgs {
    short2* nnf;    // [B][h][w]
    float* cost;    // [B][h][w]
    int w, h;
};

// d_left_right_check (bao_pmflow_refine_kernel.cu:53-76): invalid if the target is outside the image or the other
// field does not map back exactly (DIFF_THRESH 0).
__global__ void k_lr_check(short2* __restrict__ nnf, float* __restrict__ cost, const short2* __restrict__ nnf2, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const size_t off = (size_t)blockIdx.z * w * h;
    const size_t id = off + (size_t)y * w + x;
    const short2 d = nnf[id];
    bool bad;
    if (d.y < 0 || d.y >= h || d.x < 0 || d.x >= w) {
        bad = true;
    } else {
        const short2 d2 = nnf2[off + (size_t)d.y * w + d.x];
        bad = abs(d2.x - x) > 0 || abs(d2.y - y) > 0;
    }
    if (bad) {
        nnf[id] = make_short2(INVALID_LOCATION, INVALID_LOCATION);
        cost[id] = FLT_MAX;
    }
}

// d_outlier_removal (:149-182): count neighbours in (2R+1)^2 whose displacement is within +-sim of this pixel's.
// Snapshot semantics: reads `src`, writes `dst` (the reference reads and writes the same array).
__global__ void __launch_bounds__(256) k_outlier_removal(const short2* __restrict__ src, short2* __restrict__ dst, float* __restrict__ cost, int w, int h,
                                                         int R, int sim, int count_thresh) {
    extern __shared__ short2 tile[];  // (16+2R) x (16+2R) displacements (target - position); invalid marked by x = -32768
    const int TW = 16 + 2 * R;
    const size_t off = (size_t)blockIdx.z * w * h;
    const int x0 = blockIdx.x * 16 - R, y0 = blockIdx.y * 16 - R;
    for (int i = threadIdx.y * 16 + threadIdx.x; i < TW * TW; i += 256) {
        const int ty = i / TW, tx = i % TW;
        const int cx = x0 + tx, cy = y0 + ty;
        short2 v = make_short2(-32768, -32768);  // outside the image: skipped (:166)
        if (cx >= 0 && cy >= 0 && cx < w && cy < h) {
            const short2 t = src[off + (size_t)cy * w + cx];
            v = make_short2(t.x - cx, t.y - cy);  // :168-169 (also applied to invalid entries, as the reference does)
        }
        tile[i] = v;
    }
    __syncthreads();
    const int x = blockIdx.x * 16 + threadIdx.x, y = blockIdx.y * 16 + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t id = off + (size_t)y * w + x;
    const short2 cur_abs = src[id];
    short2 out = cur_abs;
    if (!(cur_abs.x < 0 && cur_abs.y < 0)) {  // :156 skip occlusion
        const short2 cur = tile[(threadIdx.y + R) * TW + threadIdx.x + R];
        int count = 0;
        for (int dy = 0; dy <= 2 * R; dy++)
            for (int dx = 0; dx <= 2 * R; dx++) {
                const int cx = x + dx - R, cy = y + dy - R;
                if (cx < 0 || cy < 0 || cx >= w || cy >= h) continue;
                const short2 nb = tile[(threadIdx.y + dy) * TW + threadIdx.x + dx];
                if (abs(nb.x - cur.x) <= sim && abs(nb.y - cur.y) <= sim) count++;
            }
        if (count < count_thresh) {
            out = make_short2(INVALID_LOCATION, INVALID_LOCATION);
            cost[id] = FLT_MAX;
        }
    }
    dst[id] = out;
}

// Compaction of the pixels d_weighted_median_filtering would process (:212-213): all pixels, or only occluded ones.
__global__ void k_collect_occluded(const short2* __restrict__ nnf, int total, bool only_occlusion, int* __restrict__ list, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool take = false;
    if (i < total) {
        const short2 d = nnf[i];
        take = !(only_occlusion && d.x >= 0 && d.y >= 0);
    }
    const unsigned m = __ballot_sync(0xffffffffu, take);
    int base = 0;
    const int lane = threadIdx.x & 31;
    if (lane == 0 && m) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (take) list[base + __popc(m & ((1u << lane) - 1))] = i;
}

// d_weighted_median_filtering (:206-259), one warp per listed pixel.
//   candidate set  : valid displacements in the (2R+1)^2 window, scanned dy-outer / dx-inner (:218)
//   candidate cost : sum over valid window pixels of w_k * max(|du|,|dv|), accumulated in window order (:231-246)
//   w_k            : __expf(-(dr*dr)/(sig_r*sig_r)) * (G[|dx|]*G[|dy|])                       (:198-204)
//   result         : first candidate with weightSum > 0 and strictly smallest cost (:248-253); untouched if none.
// The list is ordered arbitrarily (atomics) which is irrelevant: pixels are independent under snapshot semantics.
template <int R>
__global__ void __launch_bounds__(128) k_wmf(const short2* __restrict__ src, short2* __restrict__ dst, const float4* __restrict__ pix, size_t plane, int pw,
                                             int w, int h, const int* __restrict__ list, const int* __restrict__ count, float neg_sig_r2,
                                             const __grid_constant__ WmfLut lut) {
    constexpr int N = (2 * R + 1) * (2 * R + 1);
    __shared__ float s_w[4][N];
    __shared__ short2 s_d[4][N];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_list = *count;
    for (int li = blockIdx.x * 4 + warp; li < n_list; li += gridDim.x * 4) {
        const int gi = list[li];
        const int b = gi / (w * h), rem = gi - b * (w * h);
        const int y = rem / w, x = rem - y * w;
        const short2* f = src + (size_t)b * w * h;
        const float4* img = pix + (size_t)b * plane + (size_t)PAD * pw + PAD;
        const float4 c = ldpix(img + (size_t)y * pw + x);
        __syncwarp();
        for (int k = lane; k < N; k += 32) {
            const int dy = k / (2 * R + 1) - R, dx = k % (2 * R + 1) - R;
            const int cx = x + dx, cy = y + dy;
            short2 d = make_short2(-32768, 0);  // marks "skip" (outside image or invalid, :233,235)
            float wk = 0.f;
            if (cx >= 0 && cy >= 0 && cx < w && cy < h) {
                const short2 t = f[(size_t)cy * w + cx];
                if (t.x >= 0 && t.y >= 0) {
                    d = make_short2(t.x - cx, t.y - cy);
                    const float4 p = ldpix(img + (size_t)cy * pw + cx);
                    const float dr = max3abs_diff(p, c);
                    const float coef_r = __expf(__fdiv_rn(__fmul_rn(dr, dr), neg_sig_r2));  // -(dr*dr)/(sig_r*sig_r): sign folded into the divisor
                    const float coef_s = __fmul_rn(lut.g[abs(dx)], lut.g[abs(dy)]);
                    wk = __fmul_rn(coef_r, coef_s);
                }
            }
            s_w[warp][k] = wk;
            s_d[warp][k] = d;
        }
        __syncwarp();
        float best = FLT_MAX;
        int best_k = N;  // N = no candidate
        // a lane scores its (up to) NC candidates k = lane, lane + 32, ... in ONE pass over the window: the window entry (displacement, weight) is read
        // from shared memory once for all of them (the kernel was bound by those broadcast reads), and the weight sum -- the same for every
        // candidate, same taps in the same order -- is formed once.  Each candidate still adds its taps in window order.
        constexpr int NC = (N + 31) / 32;
        short2 cand[NC];
        bool has[NC];
        float cost_sum[NC];
#pragma unroll
        for (int cI = 0; cI < NC; cI++) {
            const int k = lane + 32 * cI;
            cand[cI] = k < N ? s_d[warp][k] : make_short2(-32768, 0);
            has[cI] = cand[cI].x != -32768;
            cost_sum[cI] = 0.f;
        }
        float weight_sum = 0.f;
#pragma unroll 9
        for (int q = 0; q < N; q++) {
            const short2 cur = s_d[warp][q];
            if (cur.x == -32768) continue;
            const float wq = s_w[warp][q];
#pragma unroll
            for (int cI = 0; cI < NC; cI++) {
                const int dist = max(abs(cand[cI].x - cur.x), abs(cand[cI].y - cur.y));
                cost_sum[cI] = __fmaf_rn(wq, (float)dist, cost_sum[cI]);  // :244
            }
            weight_sum = __fadd_rn(weight_sum, wq);
        }
#pragma unroll
        for (int cI = 0; cI < NC; cI++)
            if (has[cI] && weight_sum > 0.0f && cost_sum[cI] < best) {  // lane-local scan is in increasing k: first minimum wins
                best = cost_sum[cI];
                best_k = lane + 32 * cI;
            }
        // warp arg-min with the reference's order: smallest cost, ties -> smallest candidate index
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
            if (ok < N && (best_k >= N || ob < best || (ob == best && ok < best_k))) {
                best = ob;
                best_k = ok;
            }
        }
        if (lane == 0) {
            short2 out = f[(size_t)y * w + x];
            if (best_k < N) {
                const short2 cand = s_d[warp][best_k];
                out = make_short2(cand.x + x, cand.y + y);  // :251-252
            }
            // :257 the reference returns without storing when the output is still invalid; dst already holds the snapshot value
            if (out.x >= 0 && out.y >= 0) dst[(size_t)b * w * h + (size_t)y * w + x] = out;
        }
    }
}

// d_fill_holes (:297-371), snapshot semantics.  One thread per pixel (holes are sparse after the weighted median).
__global__ void k_fill_holes(const short2* __restrict__ src, short2* __restrict__ dst, const float4* __restrict__ pix, size_t plane, int pw, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const short2* f = src + (size_t)blockIdx.z * w * h;
    const size_t id = (size_t)y * w + x;
    short2 cur = f[id];
    if (cur.x >= 0 && cur.y >= 0) {
        dst[(size_t)blockIdx.z * w * h + id] = cur;
        return;
    }
    short2 nd[4] = {cur, cur, cur, cur};
    int nx[4] = {x, x, x, x}, ny[4] = {y, y, y, y};
    for (int cx = x - 1; cx >= 0; cx--) {  // left
        nd[0] = f[(size_t)y * w + cx];
        if (nd[0].x >= 0 && nd[0].y >= 0) { nx[0] = cx; break; }
    }
    for (int cx = x + 1; cx < w; cx++) {  // right
        nd[1] = f[(size_t)y * w + cx];
        if (nd[1].x >= 0 && nd[1].y >= 0) { nx[1] = cx; break; }
    }
    for (int cy = y - 1; cy >= 0; cy--) {  // up
        nd[2] = f[(size_t)cy * w + x];
        if (nd[2].x >= 0 && nd[2].y >= 0) { ny[2] = cy; break; }
    }
    for (int cy = y + 1; cy < h; cy++) {  // down
        nd[3] = f[(size_t)cy * w + x];
        if (nd[3].x >= 0 && nd[3].y >= 0) { ny[3] = cy; break; }
    }
    const float4* img = pix + (size_t)blockIdx.z * plane + (size_t)PAD * pw + PAD;
    const float4 c = ldpix(img + (size_t)y * pw + x);
    float min_diff = FLT_MAX;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float4 p = ldpix(img + (size_t)ny[i] * pw + nx[i]);
        const float diff = max3abs_diff(p, c);  // _d_rgb_max_dist (:291-295)
        if (diff < min_diff && nd[i].x >= 0 && nd[i].y >= 0) {
            min_diff = diff;
            cur.x = nd[i].x - nx[i];
            cur.y = nd[i].y - ny[i];
        }
    }
    // :368-370 the position is added back even when no direction had a valid pixel (cur still -10000,-10000)
    cur.x += x;
    cur.y += y;
    dst[(size_t)blockIdx.z * w * h + id] = cur;
}

// d_convert_nnf_to_flow (:636-655)
__global__ void k_nnf_to_flow(const short2* __restrict__ nnf, float2* __restrict__ flow, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const size_t id = (size_t)blockIdx.z * w * h + (size_t)y * w + x;
    const short2 d = nnf[id];
    float2 f;
    if (d.x <= INVALID_LOCATION || d.y <= INVALID_LOCATION) f = make_float2(EPPM_UNKNOWN_FLOW, EPPM_UNKNOWN_FLOW);
    else f = make_float2((float)(d.x - x), (float)(d.y - y));
    flow[id] = f;
}

void wmf_sweeps(eppm_context* c, short2*& cur, short2*& other, const float4* pix, size_t plane, int pw, int w, int h, int n, int iters,
                bool only_occlusion) {
    const int total = n * w * h;
    const float sr = c->prm.wmf_sig_r;
    const float neg_sig_r2 = -(sr * sr);
    for (int it = 0; it < iters; it++) {
        cudaMemsetAsync(c->occl_count, 0, sizeof(int), c->stream);
        k_collect_occluded<<<(total + 255) / 256, 256, 0, c->stream>>>(cur, total, only_occlusion, c->occl_list, c->occl_count);
        cudaMemcpyAsync(other, cur, (size_t)total * sizeof(short2), cudaMemcpyDeviceToDevice, c->stream);
        k_wmf<4><<<148 * 4, 128, 0, c->stream>>>(cur, other, pix, plane, pw, w, h, c->occl_list, c->occl_count, neg_sig_r2, c->wmf_lut);
        EPPM_LAUNCH_COUNT(2);
        short2* t = cur; cur = other; other = t;
    }
}

void op_lr_check(cudaStream_t s, short2* nnf, float* cost, const short2* nnf2, int w, int h, int n) {
    dim3 blk(128), grd((w + 127) / 128, h, n);
    k_lr_check<<<grd, blk, 0, s>>>(nnf, cost, nnf2, w, h);
    EPPM_LAUNCH_COUNT(1);
}

void op_outlier_removal(cudaStream_t s, const short2* src, short2* dst, float* cost, int w, int h, int n, int R, int sim) {
    const int thresh = (2 * R + 1) * (2 * R + 1) / 2;  // STAT_COUNT_THRESH (:146)
    dim3 b2(16, 16), g2((w + 15) / 16, (h + 15) / 16, n);
    size_t smem = (size_t)(16 + 2 * R) * (16 + 2 * R) * sizeof(short2);
    k_outlier_removal<<<g2, b2, smem, s>>>(src, dst, cost, w, h, R, sim, thresh);
    EPPM_LAUNCH_COUNT(1);
}

void op_fill_holes(cudaStream_t s, const short2* src, short2* dst, const float4* pix, size_t plane, int pw, int w, int h, int n) {
    dim3 blk(128), grd((w + 127) / 128, h, n);
    k_fill_holes<<<grd, blk, 0, s>>>(src, dst, pix, plane, pw, w, h);
    EPPM_LAUNCH_COUNT(1);
}

void op_nnf_to_flow(cudaStream_t s, const short2* nnf, float2* flow, int w, int h, int n) {
    dim3 blk(128), grd((w + 127) / 128, h, n);
    k_nnf_to_flow<<<grd, blk, 0, s>>>(nnf, flow, w, h);
    EPPM_LAUNCH_COUNT(1);
}

void run_consistency(eppm_context* c) {
    const int L = c->n_levels - 1, n = c->n_cur;
    const LevelGeom& g = c->lv[L];
    cudaStream_t s = c->stream;
    // forward then backward: the backward pass sees the forward invalidations (:78-92)
    op_lr_check(s, c->nnf[0], c->cost[0], c->nnf[1], g.w, g.h, n);
    op_lr_check(s, c->nnf[1], c->cost[1], c->nnf[0], g.w, g.h, n);
    short2* cur = c->nnf[0];
    short2* other = c->nnf_tmp;
    if (c->inplace) {   // the reference's own update order: both filters rewrite the field they are reading
        op_outlier_inplace(c, cur, c->cost[0], g.w, g.h, n);
        op_wmf_inplace(c, cur, c->pix[0][L], g.plane, g.pw, g.w, g.h, n, c->prm.wmf_iters, true);
    } else {
        op_outlier_removal(s, cur, other, c->cost[0], g.w, g.h, n, c->prm.stat_radius, c->prm.stat_sim_thresh);
        { short2* t = cur; cur = other; other = t; }
        wmf_sweeps(c, cur, other, c->pix[0][L], g.plane, g.pw, g.w, g.h, n, c->prm.wmf_iters, true);
    }
    op_fill_holes(s, cur, other, c->pix[0][L], g.plane, g.pw, g.w, g.h, n);
    { short2* t = cur; cur = other; other = t; }
    // keep the forward field where callers expect it
    if (cur != c->nnf[0]) cudaMemcpyAsync(c->nnf[0], cur, (size_t)n * g.w * g.h * sizeof(short2), cudaMemcpyDeviceToDevice, s);
    op_nnf_to_flow(s, c->nnf[0], c->flow[L], g.w, g.h, n);
}

}  // namespace eppm
