// TEST INFRASTRUCTURE ONLY — compiled into oracle/_ref/libeppm_ref.so, never into the product.
//
// Plain-C entry points around the UNMODIFIED reference host class
// (bao_flow_patchmatch_multiscale_cuda.h:33-45) so that Python tests / bench.py can drive the
// reference build through ctypes: create, init(h,w), set_data, compute_flow, plus read-outs of
// the device pyramids the class owns privately (…cuda.h:72-97) for stage-level comparisons.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define private public  // test-only: read the reference's device pyramids
#include "bao_flow_patchmatch_multiscale_cuda.h"
#undef private
#include "bao_basic.h"
#include "defs.h"
#include "bao_flow_tools.h"

struct ref_ctx {
    bao_flow_patchmatch_multiscale_cuda* eppm;
    int h, w;
    unsigned char*** img1;
    unsigned char*** img2;
    float** u;
    float** v;
    cudaEvent_t ev0, ev1;
};

extern "C" void* ref_create(int h, int w) {
    ref_ctx* c = new ref_ctx;
    c->eppm = new bao_flow_patchmatch_multiscale_cuda;
    c->h = h; c->w = w;
    c->img1 = bao_alloc<unsigned char>(h, w, 3);
    c->img2 = bao_alloc<unsigned char>(h, w, 3);
    c->u = bao_alloc<float>(h, w);
    c->v = bao_alloc<float>(h, w);
    c->eppm->init(h, w);
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    return c;
}

extern "C" void ref_destroy(void* p) {
    ref_ctx* c = (ref_ctx*)p;
    delete c->eppm;
    bao_free(c->img1); bao_free(c->img2); bao_free(c->u); bao_free(c->v);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    delete c;
}

// rgb1/rgb2: packed [h][w][3] u8 (the bao_alloc layout main.cpp:42-57 uses).
extern "C" void ref_set_data(void* p, const unsigned char* rgb1, const unsigned char* rgb2) {
    ref_ctx* c = (ref_ctx*)p;
    memcpy(c->img1[0][0], rgb1, (size_t)c->h * c->w * 3);
    memcpy(c->img2[0][0], rgb2, (size_t)c->h * c->w * 3);
    c->eppm->set_data(c->img1, c->img2);
}

// flow_uv: [h][w][2] float, interleaved (u,v).
extern "C" void ref_compute_flow(void* p, float* flow_uv) {
    ref_ctx* c = (ref_ctx*)p;
    c->eppm->compute_flow(c->u, c->v);
    const float* u = c->u[0];
    const float* v = c->v[0];
    for (size_t i = 0; i < (size_t)c->h * c->w; i++) { flow_uv[2 * i] = u[i]; flow_uv[2 * i + 1] = v[i]; }
}

// set_data + compute_flow for one pair, timed with CUDA events on the legacy default stream the
// reference launches on.  Only the two class calls are bracketed (both end in blocking copies, so the events see all device work
// plus the class's own host loops); the harness's copy of the caller's frames into bao_alloc buffers and the (u, v) interleave
// of the result happen outside the bracket.  Returns ms.
extern "C" float ref_time_pair(void* p, const unsigned char* rgb1, const unsigned char* rgb2, float* flow_uv) {
    ref_ctx* c = (ref_ctx*)p;
    memcpy(c->img1[0][0], rgb1, (size_t)c->h * c->w * 3);
    memcpy(c->img2[0][0], rgb2, (size_t)c->h * c->w * 3);
    cudaDeviceSynchronize();
    cudaEventRecord(c->ev0, 0);
    c->eppm->set_data(c->img1, c->img2);
    c->eppm->compute_flow(c->u, c->v);
    cudaEventRecord(c->ev1, 0);
    cudaEventSynchronize(c->ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    const float* u = c->u[0];
    const float* v = c->v[0];
    for (size_t i = 0; i < (size_t)c->h * c->w; i++) { flow_uv[2 * i] = u[i]; flow_uv[2 * i + 1] = v[i]; }
    return ms;
}

// texture binds issued by the shim since load (oracle/ref_shim/texref_shim.h): binds, and how many created a texture object
extern "C" unsigned long long g_texref_binds, g_texref_creates;
extern "C" void ref_shim_stats(unsigned long long* binds, unsigned long long* creates) {
    *binds = g_texref_binds;
    *creates = g_texref_creates;
}

extern "C" int ref_num_levels(void* p) { return ((ref_ctx*)p)->eppm->m_nLevels; }
extern "C" void ref_level_dims(void* p, int level, int* h, int* w) {
    ref_ctx* c = (ref_ctx*)p;
    *h = c->eppm->m_h_arr[level];
    *w = c->eppm->m_w_arr[level];
}

// which: 0 img1 pyramid (uchar4), 1 img2 pyramid (uchar4), 2 census1 (u8), 3 census2 (u8),
//        4 NNF fwd (short2), 5 NNF bwd (short2), 6 cost fwd (float), 7 cost bwd (float), 8 flow (float2).
// Copies the level plane densely (no pitch) into `out`; returns bytes written.
extern "C" long ref_read_plane(void* p, int which, int level, void* out) {
    ref_ctx* c = (ref_ctx*)p;
    bao_flow_patchmatch_multiscale_cuda* e = c->eppm;
    int h = e->m_h_arr[level], w = e->m_w_arr[level];
    switch (which) {
    case 0: case 1: {
        uchar4* src = (which == 0 ? e->m_img1_pyr : e->m_img2_pyr)[level];
        cudaMemcpy2D(out, (size_t)w * 4, src, e->m_arrPitchUchar4[level], (size_t)w * 4, h, cudaMemcpyDeviceToHost);
        return (long)w * h * 4;
    }
    case 2: case 3: {
        unsigned char* src = (which == 2 ? e->m_img1_census_pyramid : e->m_img2_census_pyramid)[level];
        cudaMemcpy2D(out, (size_t)w, src, e->m_arrPitchUchar1[level], (size_t)w, h, cudaMemcpyDeviceToHost);
        return (long)w * h;
    }
    case 4: case 5: {
        short2* src = (which == 4 ? e->m_disp_vec1_pyramid : e->m_disp_vec2_pyramid)[level];
        cudaMemcpy(out, src, (size_t)w * h * 4, cudaMemcpyDeviceToHost);
        return (long)w * h * 4;
    }
    case 6: case 7: {
        float* src = (which == 6 ? e->m_pmcost1_pyramid : e->m_pmcost2_pyramid)[level];
        cudaMemcpy(out, src, (size_t)w * h * 4, cudaMemcpyDeviceToHost);
        return (long)w * h * 4;
    }
    case 8: {
        cudaMemcpy(out, e->m_flow1_pyramid[level], (size_t)w * h * 8, cudaMemcpyDeviceToHost);
        return (long)w * h * 8;
    }
    }
    return -1;
}

// The reference's own evaluation and .flo writer (basic/bao_flow_tools.cpp:49-141) on interleaved (u, v) arrays, for the tests of
// eppm_eval_flow / eppm_write_flo.
extern "C" void ref_calc_flow_error(const float* flow_uv, const float* gt_uv, int h, int w, int border, int error_thresh, float* epe, float* aae, float* outlier_frac) {
    float** u = bao_alloc<float>(h, w); float** v = bao_alloc<float>(h, w);
    float** gu = bao_alloc<float>(h, w); float** gv = bao_alloc<float>(h, w);
    for (size_t i = 0; i < (size_t)h * w; i++) { u[0][i] = flow_uv[2 * i]; v[0][i] = flow_uv[2 * i + 1]; gu[0][i] = gt_uv[2 * i]; gv[0][i] = gt_uv[2 * i + 1]; }
    *epe = 0.f; *aae = 0.f;
    bao_calc_flow_error(u, v, gu, gv, h, w, *epe, *aae, border, false);
    *outlier_frac = bao_calc_flow_error_percentage(u, v, gu, gv, h, w, error_thresh, NULL);
    bao_free(u); bao_free(v); bao_free(gu); bao_free(gv);
}
extern "C" void ref_save_flo(const char* path, const float* flow_uv, int h, int w) {
    float** u = bao_alloc<float>(h, w); float** v = bao_alloc<float>(h, w);
    for (size_t i = 0; i < (size_t)h * w; i++) { u[0][i] = flow_uv[2 * i]; v[0][i] = flow_uv[2 * i + 1]; }
    bao_save_flo_file(path, u, v, h, w);
    bao_free(u); bao_free(v);
}
