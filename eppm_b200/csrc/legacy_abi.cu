// The reference's stage functions (include/eppm_legacy_abi.h) on FOREIGN buffers in the reference's layouts:
// uchar4 / u8 planes with a byte pitch, dense short2 / float / float2 planes.  Each entry point converts the foreign
// image planes into this library's packed planes (one small kernel), runs the same kernels the eppm_* API runs, and
// writes the result back in the foreign layout.  A context per (h, w) is created on first use and cached.
#include <float.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <tuple>
#include <utility>

#include "../../include/eppm_legacy_abi.h"
#include "eppm_internal.h"

using namespace eppm;

namespace {

std::mutex g_mu;
// One legacy call at a time: like the reference (file-scope textures and __constant__ tables, one global RNG state) the stage functions are
// not re-entrant -- two host threads would otherwise share one cached context.  Held by Scope for the duration of a call.
std::recursive_mutex g_call_mu;
constexpr size_t MAX_CACHED = 8;   // contexts per cache; the least recently used one is destroyed beyond that (a caller cycling through frame sizes)
unsigned long long g_use_clock = 0;
std::map<eppm_context*, unsigned long long> g_last_use;
typedef std::map<std::tuple<int, int, int>, eppm_context*> CtxCache;   // keyed by (device, h, w)
CtxCache g_single;    // single-level contexts
CtxCache g_pyramid;   // full-pyramid contexts, (h, w) of level 0

void complain(const char* where) { fprintf(stderr, "EPPM(b200) %s: %s\n", where, eppm_last_error()); }

// Returns with g_call_mu HELD when it returns a context (the caller's Scope adopts and releases it): the context cannot be evicted or
// used by another thread between the lookup and the end of the call.
eppm_context* get_ctx(CtxCache& cache, int h, int w, int levels) {
    g_call_mu.lock();
    std::lock_guard<std::mutex> lk(g_mu);
    int dev = 0;
    cudaGetDevice(&dev);   // the caller's current device, like the reference's implicit context
    const std::tuple<int, int, int> key(dev, h, w);
    auto it = cache.find(key);
    if (it != cache.end() && it->second->n_levels == levels) { g_last_use[it->second] = ++g_use_clock; return it->second; }
    if (it != cache.end()) { g_last_use.erase(it->second); eppm_destroy(it->second); cache.erase(it); }
    while (cache.size() >= MAX_CACHED) {   // evict the least recently used context
        auto lru = cache.begin();
        for (auto jt = cache.begin(); jt != cache.end(); ++jt)
            if (g_last_use[jt->second] < g_last_use[lru->second]) lru = jt;
        g_last_use.erase(lru->second);
        eppm_destroy(lru->second);
        cache.erase(lru);
    }
    eppm_params p;
    eppm_default_params(&p);
    p.pyr_levels = levels;
    eppm_context* c = nullptr;
    if (eppm_create(&c, dev, h, w, 1, &p) != EPPM_OK) { complain("context"); g_call_mu.unlock(); return nullptr; }
    c->n_cur = 1;
    cache[key] = c;
    g_last_use[c] = ++g_use_clock;
    return c;
}

// entry: make work queued on the legacy default stream visible; exit: our stream is drained
struct Scope {
    eppm_context* c;
    const char* name;
    std::unique_lock<std::recursive_mutex> lk;
    Scope(eppm_context* c_, const char* n) : c(c_), name(n), lk(g_call_mu, std::adopt_lock) { cudaStreamSynchronize(0); }
    ~Scope() {
        if (c && !cuda_ok(cudaStreamSynchronize(c->stream), name)) complain(name);
    }
};

template <class T>
void copy_in(eppm_context* c, T* dense, const T* foreign, size_t pitch, int w, int h) {
    cudaMemcpy2DAsync(dense, (size_t)w * sizeof(T), foreign, pitch, (size_t)w * sizeof(T), h, cudaMemcpyDeviceToDevice, c->stream);
}
template <class T>
void copy_out(eppm_context* c, T* foreign, size_t pitch, const T* dense, int w, int h) {
    cudaMemcpy2DAsync(foreign, pitch, dense, (size_t)w * sizeof(T), (size_t)w * sizeof(T), h, cudaMemcpyDeviceToDevice, c->stream);
}

// ---- stage functions the reference declares but compute_flow never calls (…cuda.cpp:40-62): small per-pixel kernels working
// directly on the caller's (pitched) buffers

// d_left_right_check_buffered (bao_pmflow_refine_kernel.cu:93-122): like the in-place check but with the looser threshold 50 and
// the result written to a second buffer
__global__ void k_lr_check_buffered(short2* __restrict__ out_nnf, float* __restrict__ out_cost, const short2* nnf, const float* cost,
                                    const short2* __restrict__ nnf2, int w, int h, size_t cost_w, size_t disp_w) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const short2 d = nnf[y * disp_w + x];
    short2 o = make_short2((short)INVALID_LOCATION, (short)INVALID_LOCATION);
    float oc = FLT_MAX;
    if (!(d.y < 0 || d.y >= h || d.x < 0 || d.x >= w)) {
        const short2 d2 = nnf2[d.y * disp_w + d.x];
        if (!(abs(d2.x - (short)x) > 50 || abs(d2.y - (short)y) > 50)) {   // DIFF_THRESH_2 (:93)
            o = d;
            oc = cost[y * cost_w + x];
        }
    }
    out_nnf[y * disp_w + x] = o;
    out_cost[y * cost_w + x] = oc;
}

// d_convert_flow_to_nnf (bao_pmflow_refine_kernel.cu:657-676)
__global__ void k_flow_to_nnf(short2* __restrict__ nnf, const float2* __restrict__ flow, int w, int h, size_t flow_w, size_t disp_w) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float2 f = flow[y * flow_w + x];
    short2 d;
    if (f.x > EPPM_UNKNOWN_FLOW_THRESH || f.y > EPPM_UNKNOWN_FLOW_THRESH) {
        d = make_short2((short)INVALID_LOCATION, (short)INVALID_LOCATION);
    } else {
        d.x = short(__fadd_rn(f.x, (float)x));   // short(curFlow.x+id_x)
        d.y = short(__fadd_rn(f.y, (float)y));
    }
    nnf[y * disp_w + x] = d;
}

// d_flow_cutoff (bao_pmflow_refine_kernel.cu:891-900) with the reference's __min/__max macros (basic/bao_basic_cuda.h:44-45)
__global__ void k_flow_cutoff(float2* __restrict__ flow, int w, int h, size_t flow_w, float m) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float2 f = flow[y * flow_w + x];
    const float nm = -m;
    const float ax = (m < f.x) ? m : f.x, ay = (m < f.y) ? m : f.y;
    f.x = (nm > ax) ? nm : ax;
    f.y = (nm > ay) ? nm : ay;
    flow[y * flow_w + x] = f;
}

// d_eliminate_still_region_flow + _d_compute_patch_dist_ad_L2 (bao_pmflow_kernel.cu:555-586, 2071-2081): unweighted mean of the
// AD term over the 100 stride-2 samples of the patch at ZERO displacement; flow := 0 where it is <= 0.1
__global__ void k_eliminate_still(float2* __restrict__ flow, const float4* __restrict__ A, const float4* __restrict__ B, int pw, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const unsigned o = (unsigned)(y + PAD) * pw + x + PAD;
    float cs = 0.f, ws = 0.f;
    for (int i = -PATCH_R; i <= PATCH_R; i += 2)
        for (int j = -PATCH_R; j <= PATCH_R; j += 2) {
            const float4 p1 = ldpix(A + (o + i * pw + j)), p2 = ldpix(B + (o + i * pw + j));
            const float c = max3abs_diff(p1, p2);
            cs = __fadd_rn(cs, exp_ad_cost(c));
            ws = __fadd_rn(ws, 1.0f);
        }
    const float cost = __fdiv_rn(cs, ws);
    if ((double)cost <= 0.1) flow[(size_t)y * w + x] = make_float2(0.f, 0.f);   // SIMILAR_MIN_COST is a double literal (:2071)
}

}  // namespace

extern "C" {

void baoCudaPatchMatchMultiscalePrepare(uchar4** pImgPyr1, uchar4** pImgPyr2, unsigned char** pCensusPyr1, unsigned char** pCensusPyr2,
                                        uchar4** pTempPyr1, uchar4** pTempPyr2, int* arrH, int* arrW, size_t* arrPitchUchar4,
                                        size_t* arrPitchUchar1, int nLevels, uchar4* d_img1, uchar4* d_img2, int h, int w) {
    (void)pTempPyr1; (void)pTempPyr2;
    eppm_context* c = get_ctx(g_pyramid, h, w, nLevels);
    if (!c) return;
    Scope sc(c, "baoCudaPatchMatchMultiscalePrepare");
    op_preblur_rgba(c, d_img1, d_img2, arrPitchUchar4[0]);
    op_pyramid_and_pack(c, 1);
    for (int i = 0; i < nLevels; i++) {
        const LevelGeom& g = c->lv[i];
        if (g.w != arrW[i] || g.h != arrH[i]) { fprintf(stderr, "EPPM(b200): level %d geometry mismatch\n", i); return; }
        copy_out(c, pImgPyr1[i], arrPitchUchar4[i], c->rgba[0][i], g.w, g.h);
        copy_out(c, pImgPyr2[i], arrPitchUchar4[i], c->rgba[1][i], g.w, g.h);
        op_extract_census(c->stream, c->pix[0][i], g, pCensusPyr1[i], arrPitchUchar1[i]);
        op_extract_census(c->stream, c->pix[1][i], g, pCensusPyr2[i], arrPitchUchar1[i]);
    }
}

void baoCudaCensusTransform(unsigned char* d_census1, unsigned char* d_census2, uchar4* d_img1, uchar4* d_img2, int w, int h, size_t img_pitch,
                            size_t census_pitch) {
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaCensusTransform");
    const LevelGeom& g = c->lv[0];
    k_pack_planes(c->stream, d_img1, img_pitch, 0, c->pix[0][0], g, 1);
    k_pack_planes(c->stream, d_img2, img_pitch, 0, c->pix[1][0], g, 1);
    op_extract_census(c->stream, c->pix[0][0], g, d_census1, census_pitch);
    op_extract_census(c->stream, c->pix[1][0], g, d_census2, census_pitch);
}

void baoCudaPatchMatch(short2* d_disp_vec, float* d_cost, uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1, unsigned char* d_census2,
                       int w, int h, size_t img_pitch, size_t cost_pitch, size_t disp_pitch, size_t census_pitch) {
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaPatchMatch");
    const LevelGeom& g = c->lv[0];
    op_pack_foreign(c->stream, d_img1, img_pitch, d_census1, census_pitch, c->pix[0][0], g);
    op_pack_foreign(c->stream, d_img2, img_pitch, d_census2, census_pitch, c->pix[1][0], g);
    for (int img = 0; img < 2; img++) op_transpose_plane(c->stream, c->pix[img][0], c->pixT[img], g, 1);
    if (c->pixQ[0])
        for (int img = 0; img < 2; img++) op_split_plane(c->stream, c->pix[img][0], c->pixQ[img], g, 1);
    run_patchmatch_dirs(c, 1);
    copy_out(c, d_disp_vec, disp_pitch, c->nnf[0], w, h);
    copy_out(c, d_cost, cost_pitch, c->cost[0], w, h);
}

// Declared by the reference's host class, called nowhere, and unfinished upstream (its forward row pass stores the candidate's scale into
// the cost plane, bao_pmflow_kernel.cu:1207; its cost ignores the census planes it is given).  Mirrored as it stands (run_patchmatch_scaled),
// bit-exact against the reference build.  Like the reference's random-field kernel (:151) it assumes scale_pitch == disp_pitch.
void baoCudaPatchMatch_Scaled(short2* d_disp_vec, float* d_scale, float* d_cost, uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1,
                              unsigned char* d_census2, int w, int h, size_t img_pitch, size_t cost_pitch, size_t disp_pitch, size_t scale_pitch,
                              size_t census_pitch) {
    if (!d_disp_vec || !d_scale || !d_cost || !d_img1 || !d_img2 || w < 1 || h < 1 || scale_pitch != disp_pitch) {
        set_error("baoCudaPatchMatch_Scaled: null plane, empty frame, or scale_pitch != disp_pitch (the reference indexes the scale plane with the displacement pitch)");
        complain("baoCudaPatchMatch_Scaled");
        return;
    }
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaPatchMatch_Scaled");
    const LevelGeom& g = c->lv[0];
    // the census planes are optional here: the scaled cost never reads them (:596-609)
    if (d_census1 && d_census2) {
        op_pack_foreign(c->stream, d_img1, img_pitch, d_census1, census_pitch, c->pix[0][0], g);
        op_pack_foreign(c->stream, d_img2, img_pitch, d_census2, census_pitch, c->pix[1][0], g);
    } else {
        k_pack_planes(c->stream, d_img1, img_pitch, 0, c->pix[0][0], g, 1);
        k_pack_planes(c->stream, d_img2, img_pitch, 0, c->pix[1][0], g, 1);
    }
    float* scale = nullptr;
    if (cudaMalloc((void**)&scale, (size_t)w * h * sizeof(float)) != cudaSuccess) { set_error("baoCudaPatchMatch_Scaled: out of device memory"); complain("baoCudaPatchMatch_Scaled"); return; }
    if (!run_patchmatch_scaled(c, scale)) { cudaFree(scale); complain("baoCudaPatchMatch_Scaled"); return; }
    copy_out(c, d_disp_vec, disp_pitch, c->nnf[0], w, h);
    copy_out(c, d_scale, scale_pitch, scale, w, h);
    copy_out(c, d_cost, cost_pitch, c->cost[0], w, h);
    cudaStreamSynchronize(c->stream);
    cudaFree(scale);
}

void baoCudaPatchMatch_PlaneFitting(short2* d_disp_vec, float* d_cost, uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1,
                                    unsigned char* d_census2, int w, int h, size_t img_pitch, size_t cost_pitch, size_t disp_pitch, size_t census_pitch) {
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaPatchMatch_PlaneFitting");
    const LevelGeom& g = c->lv[0];
    op_pack_foreign(c->stream, d_img1, img_pitch, d_census1, census_pitch, c->pix[0][0], g);
    op_pack_foreign(c->stream, d_img2, img_pitch, d_census2, census_pitch, c->pix[1][0], g);
    if (!run_patchmatch_planefitting(c)) { complain("baoCudaPatchMatch_PlaneFitting"); return; }
    copy_out(c, d_disp_vec, disp_pitch, c->nnf[0], w, h);
    copy_out(c, d_cost, cost_pitch, c->cost[0], w, h);
}

void baoCudaLeftRightCheck(short2* d_disp_vec, float* d_cost, short2* d_disp_vec2, float* d_cost2, int w, int h, size_t cost_pitch,
                           size_t disp_pitch) {
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaLeftRightCheck");
    copy_in(c, c->nnf[0], d_disp_vec, disp_pitch, w, h);
    copy_in(c, c->nnf[1], d_disp_vec2, disp_pitch, w, h);
    copy_in(c, c->cost[0], d_cost, cost_pitch, w, h);
    copy_in(c, c->cost[1], d_cost2, cost_pitch, w, h);
    op_lr_check(c->stream, c->nnf[0], c->cost[0], c->nnf[1], w, h, 1);
    op_lr_check(c->stream, c->nnf[1], c->cost[1], c->nnf[0], w, h, 1);
    copy_out(c, d_disp_vec, disp_pitch, c->nnf[0], w, h);
    copy_out(c, d_disp_vec2, disp_pitch, c->nnf[1], w, h);
    copy_out(c, d_cost, cost_pitch, c->cost[0], w, h);
    copy_out(c, d_cost2, cost_pitch, c->cost[1], w, h);
}

void baoCudaOutlierRemoval(short2* d_disp_vec, float* d_cost, int w, int h, size_t cost_pitch, size_t disp_pitch) {
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaOutlierRemoval");
    copy_in(c, c->nnf[0], d_disp_vec, disp_pitch, w, h);
    copy_in(c, c->cost[0], d_cost, cost_pitch, w, h);
    if (c->inplace) {
        op_outlier_inplace(c, c->nnf[0], c->cost[0], w, h, 1);
        copy_out(c, d_disp_vec, disp_pitch, c->nnf[0], w, h);
    } else {
        op_outlier_removal(c->stream, c->nnf[0], c->nnf_tmp, c->cost[0], w, h, 1, c->prm.stat_radius, c->prm.stat_sim_thresh);
        copy_out(c, d_disp_vec, disp_pitch, c->nnf_tmp, w, h);
    }
    copy_out(c, d_cost, cost_pitch, c->cost[0], w, h);
}

void baoCudaWeightedMedianFilter(short2* d_disp_vec, float* d_cost, uchar4* d_img, int w, int h, size_t img_pitch, size_t cost_pitch,
                                 size_t disp_pitch, int num_iter, bool is_only_occlusion) {
    (void)d_cost; (void)cost_pitch;
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaWeightedMedianFilter");
    const LevelGeom& g = c->lv[0];
    op_pack_foreign(c->stream, d_img, img_pitch, nullptr, 0, c->pix[0][0], g);
    copy_in(c, c->nnf[0], d_disp_vec, disp_pitch, w, h);
    short2* cur = c->nnf[0];
    short2* other = c->nnf_tmp;
    if (c->inplace) op_wmf_inplace(c, cur, c->pix[0][0], g.plane, g.pw, w, h, 1, num_iter, is_only_occlusion);
    else wmf_sweeps(c, cur, other, c->pix[0][0], g.plane, g.pw, w, h, 1, num_iter, is_only_occlusion);
    copy_out(c, d_disp_vec, disp_pitch, cur, w, h);
}

void baoCudaFillHole(short2* d_disp_vec, float* d_cost, uchar4* d_img, int w, int h, size_t img_pitch, size_t cost_pitch, size_t disp_pitch) {
    (void)d_cost; (void)cost_pitch;
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaFillHole");
    const LevelGeom& g = c->lv[0];
    op_pack_foreign(c->stream, d_img, img_pitch, nullptr, 0, c->pix[0][0], g);
    copy_in(c, c->nnf[0], d_disp_vec, disp_pitch, w, h);
    op_fill_holes(c->stream, c->nnf[0], c->nnf_tmp, c->pix[0][0], g.plane, g.pw, w, h, 1);
    copy_out(c, d_disp_vec, disp_pitch, c->nnf_tmp, w, h);
}

void baoCudaNNF2Flow(float2* d_flow, short2* d_disp_vec, int w, int h, size_t disp_pitch, size_t flow_pitch) {
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaNNF2Flow");
    copy_in(c, c->nnf[0], d_disp_vec, disp_pitch, w, h);
    op_nnf_to_flow(c->stream, c->nnf[0], c->flow[0], w, h, 1);
    copy_out(c, d_flow, flow_pitch, c->flow[0], w, h);
}

void baoCudaBLFCostFilterRefine(float2* d_flow_vec, uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1, unsigned char* d_census2, int w,
                                int h, size_t img_pitch, size_t census_pitch) {
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaBLFCostFilterRefine");
    const LevelGeom& g = c->lv[0];
    op_pack_foreign(c->stream, d_img1, img_pitch, d_census1, census_pitch, c->pix[0][0], g);
    op_pack_foreign(c->stream, d_img2, img_pitch, d_census2, census_pitch, c->pix[1][0], g);
    op_refine(c, c->pix[0][0], c->pix[1][0], g, d_flow_vec, w, h, 0, c->flow_tmp, 1);
    cudaMemcpyAsync(d_flow_vec, c->flow_tmp, (size_t)w * h * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream);
}

void baoCudaBLF_C2F(float2** pFlowPyr, uchar4** pImgPyr1, uchar4** pImgPyr2, unsigned char** pCensusPyr1, unsigned char** pCensusPyr2,
                    float2** pTempPyr1, float2** pTempPyr2, int* arrH, int* arrW, size_t* arrPitchUchar4, size_t* arrPitchUchar1, int nLayerIdx) {
    (void)pTempPyr1; (void)pTempPyr2;
    const int l = nLayerIdx, w = arrW[l], h = arrH[l];
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaBLF_C2F");
    const LevelGeom& g = c->lv[0];
    op_pack_foreign(c->stream, pImgPyr1[l], arrPitchUchar4[l], pCensusPyr1[l], arrPitchUchar1[l], c->pix[0][0], g);
    op_pack_foreign(c->stream, pImgPyr2[l], arrPitchUchar4[l], pCensusPyr2[l], arrPitchUchar1[l], c->pix[1][0], g);
    op_refine(c, c->pix[0][0], c->pix[1][0], g, pFlowPyr[l + 1], arrW[l + 1], arrH[l + 1], 1, pFlowPyr[l], 1);
}

void baoCudaImageSmoothing(uchar4* d_img_smoothed, uchar4* d_img, int w, int h, size_t img_pitch) {
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaImageSmoothing");
    const LevelGeom& g = c->lv[0];
    op_pack_foreign(c->stream, d_img, img_pitch, nullptr, 0, c->pix[0][0], g);
    op_image_bilateral(c, d_img_smoothed, d_img, img_pitch, c->pix[0][0], g);
}

void baoCudaFlowBilteralUpsampling(float2* d_flow_vec, uchar4* d_img, int w, int h, size_t img_pitch, float2* d_flow_vec_small, int w_s, int h_s,
                                   float ratio_up) {
    (void)h_s;
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaFlowBilteralUpsampling");
    const LevelGeom& g = c->lv[0];
    op_pack_foreign(c->stream, d_img, img_pitch, nullptr, 0, c->pix[0][0], g);
    op_flow_bilateral_upsample(c, d_flow_vec, c->pix[0][0], g, d_flow_vec_small, w_s, ratio_up);   // both flow planes dense (:883-884)
}

void baoCudaFlowSmoothing(float2* d_flow, uchar4* d_img, int w, int h, size_t img_pitch, size_t flow_pitch) {
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoCudaFlowSmoothing");
    const LevelGeom& g = c->lv[0];
    op_pack_foreign(c->stream, d_img, img_pitch, nullptr, 0, c->pix[0][0], g);
    copy_in(c, c->flow[0], d_flow, flow_pitch, w, h);
    if (c->inplace) {
        op_smooth_inplace(c, c->flow[0], c->pix[0][0], g, 1);
        copy_out(c, d_flow, flow_pitch, c->flow[0], w, h);
    } else {
        op_smooth(c, c->flow[0], c->flow_tmp, c->pix[0][0], g, 1);
        copy_out(c, d_flow, flow_pitch, c->flow_tmp, w, h);
    }
}

void baoCudaLeftRightCheck_Buffered(short2* d_disp_vec, float* d_cost, short2* d_disp_vec2, float* d_cost2, short2* d_disp_vec_temp,
                                    float* d_cost_temp, int w, int h, size_t cost_pitch, size_t disp_pitch) {
    // bao_pmflow_refine_kernel.cu:124-140: forward into the temp buffers, backward in place (a pixel only rewrites itself), then the
    // temp buffers are copied back as DENSE h*w arrays (bao_cuda_copy_d2d) -- exact only for dense planes, like the reference
    cudaStreamSynchronize(0);
    const size_t cw = cost_pitch / sizeof(float), dw = disp_pitch / sizeof(short2);
    dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
    k_lr_check_buffered<<<grd, blk>>>(d_disp_vec_temp, d_cost_temp, d_disp_vec, d_cost, d_disp_vec2, w, h, cw, dw);
    k_lr_check_buffered<<<grd, blk>>>(d_disp_vec2, d_cost2, d_disp_vec2, d_cost2, d_disp_vec, w, h, cw, dw);
    cudaMemcpyAsync(d_disp_vec, d_disp_vec_temp, sizeof(short2) * (size_t)w * h, cudaMemcpyDeviceToDevice, 0);
    cudaMemcpyAsync(d_cost, d_cost_temp, sizeof(float) * (size_t)w * h, cudaMemcpyDeviceToDevice, 0);
    EPPM_LAUNCH_COUNT(2);
    if (!cuda_ok(cudaStreamSynchronize(0), "baoCudaLeftRightCheck_Buffered")) complain("baoCudaLeftRightCheck_Buffered");
}

void baoCudaFlow2NNF(short2* d_disp_vec, float2* d_flow, int w, int h, size_t disp_pitch, size_t flow_pitch) {
    cudaStreamSynchronize(0);
    dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
    k_flow_to_nnf<<<grd, blk>>>(d_disp_vec, d_flow, w, h, flow_pitch / sizeof(float2), disp_pitch / sizeof(short2));
    EPPM_LAUNCH_COUNT(1);
    if (!cuda_ok(cudaStreamSynchronize(0), "baoCudaFlow2NNF")) complain("baoCudaFlow2NNF");
}

void baoCudaFlowCutoff(float2* d_flow, int w, int h, size_t flow_pitch, float max_flow_val) {
    cudaStreamSynchronize(0);
    dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
    k_flow_cutoff<<<grd, blk>>>(d_flow, w, h, flow_pitch / sizeof(float2), max_flow_val);
    EPPM_LAUNCH_COUNT(1);
    if (!cuda_ok(cudaStreamSynchronize(0), "baoCudaFlowCutoff")) complain("baoCudaFlowCutoff");
}

void baoEliminateStillRegionFlow(float2* d_flow, uchar4* d_img1, uchar4* d_img2, int w, int h, size_t img_pitch) {
    eppm_context* c = get_ctx(g_single, h, w, 1);
    if (!c) return;
    Scope sc(c, "baoEliminateStillRegionFlow");
    const LevelGeom& g = c->lv[0];
    op_pack_foreign(c->stream, d_img1, img_pitch, nullptr, 0, c->pix[0][0], g);
    op_pack_foreign(c->stream, d_img2, img_pitch, nullptr, 0, c->pix[1][0], g);
    dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
    k_eliminate_still<<<grd, blk, 0, c->stream>>>(d_flow, c->pix[0][0], c->pix[1][0], g.pw, w, h);   // the flow plane is dense (:2090)
    EPPM_LAUNCH_COUNT(1);
}

}  // extern "C"

// ---- C++ linkage, like the reference: the one non-extern-"C" symbol its host class needs (…cuda.cpp:64, :311) ----
// bao_cuda_convert_flow_to_colorshow (basic/bao_basic_cuda.cuh:745-849): Middlebury colour wheel on the device.  Visualisation only;
// arithmetic follows the reference's expressions (double where its literals are double), libdevice atan2f like the reference.
namespace {
__constant__ unsigned char c_wheel[55][3];

__global__ void k_flow_to_color(uchar4* __restrict__ rgb, const float2* __restrict__ flow, int h, int w, float max_rad) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float2 v = flow[(size_t)y * w + x];
    uchar4 o = make_uchar4(0, 0, 0, 0);
    if (fabsf(v.x) < 999999.f && fabsf(v.y) < 999999.f) {                  // :823 unknown flow stays black
        const float fx = v.x / max_rad, fy = v.y / max_rad;
        const float rad = __fsqrt_rn(fx * fx + fy * fy);                  // :780
        const float a = atan2f(-fy, -fx) / 3.14159f;
        const float fk = (a + 1.0f) / 2.0f * (55 - 1);
        const int k0 = (int)fk, k1 = (k0 + 1) % 55;
        const float f = fk - k0;
        unsigned char pv[3];
        for (int b = 0; b < 3; b++) {
            const float col0 = c_wheel[k0][b] / 255.0f, col1 = c_wheel[k1][b] / 255.0f;
            float col = (1 - f) * col0 + f * col1;
            if (rad <= 1) col = 1 - rad * (1 - col);                       // increase saturation with radius
            else col *= .75;                                               // out of range (double literal in the reference)
            pv[b] = (unsigned char)(int)(255.0 * col);
        }
        o = make_uchar4(pv[0], pv[1], pv[2], 0);
    }
    rgb[(size_t)y * w + x] = o;
}
}  // namespace

void bao_cuda_convert_flow_to_colorshow(uchar4* rgbflow, float2* flow_vec, int h, int w, float max_disp_x, float max_disp_y) {
    cudaStreamSynchronize(0);
    {   // 165 bytes of __constant__ per call: constant memory is per device, a process may use several
        unsigned char wh[55][3];
        const int RY = 15, YG = 6, GC = 4, CB = 11, BM = 13, MR = 6;     // :760-773
        int k = 0;
        for (int i = 0; i < RY; i++, k++) { wh[k][0] = 255; wh[k][1] = 255 * i / RY; wh[k][2] = 0; }
        for (int i = 0; i < YG; i++, k++) { wh[k][0] = 255 - 255 * i / YG; wh[k][1] = 255; wh[k][2] = 0; }
        for (int i = 0; i < GC; i++, k++) { wh[k][0] = 0; wh[k][1] = 255; wh[k][2] = 255 * i / GC; }
        for (int i = 0; i < CB; i++, k++) { wh[k][0] = 0; wh[k][1] = 255 - 255 * i / CB; wh[k][2] = 255; }
        for (int i = 0; i < BM; i++, k++) { wh[k][0] = 255 * i / BM; wh[k][1] = 0; wh[k][2] = 255; }
        for (int i = 0; i < MR; i++, k++) { wh[k][0] = 255; wh[k][1] = 0; wh[k][2] = 255 - 255 * i / MR; }
        cudaMemcpyToSymbol(c_wheel, wh, sizeof(wh));
    }
    const float max_rad = sqrt(max_disp_x * max_disp_x + max_disp_y * max_disp_y);   // :841
    k_flow_to_color<<<dim3((w + 31) / 32, (h + 7) / 8), dim3(32, 8)>>>(rgbflow, flow_vec, h, w, max_rad);
    EPPM_LAUNCH_COUNT(1);
    if (!cuda_ok(cudaStreamSynchronize(0), "bao_cuda_convert_flow_to_colorshow")) complain("bao_cuda_convert_flow_to_colorshow");
}

