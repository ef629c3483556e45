// Context, arena and the public eppm_* entry points (include/eppm.h).
// Host sequencing restates bao_flow_patchmatch_multiscale_cuda::init/set_data/compute_flow
// (bao_flow_patchmatch_multiscale_cuda.cpp:112-168,217-306) for a whole batch of pairs per call.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "eppm_internal.h"

namespace eppm {

std::atomic<unsigned long long> g_launches{0};
static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
bool cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return false;
}

__global__ void k_extract_census(const float4* __restrict__ pix, int pw, unsigned char* __restrict__ out, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    out[(size_t)y * w + x] = (unsigned char)unpack_census(pix[(size_t)(y + PAD) * pw + x + PAD].w);
}

static void host_luts(eppm_context* c) {
    // _initGaussianLookupTable (bao_pmflow_kernel.cu:670-687), evaluated with the host libm like the reference does.
    // `volatile` keeps the compiler from folding expf at build time: the reference calls it at run time.
    const eppm_params& p = c->prm;
    volatile float sig_s = 0.5f * p.patch_r;  // PM_SIG_S (defs.h:45)
    float G[10];
    for (int i = 0; i <= 9; i++) G[i] = i <= p.patch_r ? expf(-(i * i) / (sig_s * sig_s)) : 0.f;
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 10; j++) c->cost_lut.gg[i][j] = G[j] * G[i];  // cSpatialGaussian[abs(j)] * cSpatialGaussian[abs(i)] (:293)
    volatile float lc = p.lambda_census;
    for (int i = 0; i <= 8; i++) c->cost_lut.census[i] = 1 - expf(-float(i * i) / (lc * 8 * lc * 8));  // :683
    volatile float wmf_sig_s = p.wmf_radius * 1.0f;  // WMF_SIG_S (defs.h:59)
    for (int i = 0; i < 8; i++) c->wmf_lut.g[i] = i <= p.wmf_radius ? expf(-float(i * i) / (wmf_sig_s * wmf_sig_s)) : 0.f;  // refine:273
    volatile int bs = p.blf_sig_s;
    for (int i = 0; i < 21; i++) c->smooth_lut.g[i] = i <= 2 * p.blf_sig_s ? expf(-float(i * i) / float(bs * bs)) : 0.f;  // refine:812
}

bool ensure_pm_buffers(eppm_context* c) {
    if (c->pm_arena.base) return true;
    const eppm_params& p = c->prm;
    const LevelGeom& gc = c->lv[c->n_levels - 1];
    const size_t B = c->max_batch, nc = (size_t)gc.w * gc.h;
    for (int pass = 0; pass < 2; pass++) {
        Arena& A = c->pm_arena;
        A.used = 0;
        if (pass == 1 && !cuda_ok(cudaMalloc((void**)&A.base, A.size), "cudaMalloc(PatchMatch arena)")) { A.base = nullptr; return false; }
        const int sl = p.prop_seg_length;
        // chains of a pass: scan lines (rounded up to whole warps) x segments per line
        const size_t thr_row = (size_t)((gc.h + 31) & ~31) * ((gc.w + sl - 1) / sl), thr_col = (size_t)((gc.w + 31) & ~31) * ((gc.h + sl - 1) / sl);
        const size_t thr = 2 * B * (thr_row > thr_col ? thr_row : thr_col);
        c->prop_prev = A.take<short2>(thr);
        c->prop_queue = A.take<int4>(thr);
        c->prop_memo = A.take<int4>(2 * B * nc);
        c->prop_count = A.take<int>((size_t)(p.num_iter > 0 ? p.num_iter : 1) * 4 * sl);
        c->rng_init = A.take<short2>(nc);
        c->rng_search = A.take<short2>((size_t)(p.num_iter > 0 ? p.num_iter : 1) * (p.num_rand_guess > 0 ? p.num_rand_guess : 1) * nc);
        if (pass == 0) A.size = A.used;
    }
    cudaMemsetAsync(c->prop_memo, 0xff, (size_t)2 * B * nc * sizeof(int4), c->stream);
    return true;
}

// Every entry point runs on the context's device and puts the caller's current device back afterwards (a process that drives several
// GPUs from one thread, or Python's garbage collector calling eppm_destroy at an arbitrary moment, must not find its device changed).
struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace eppm

using namespace eppm;

extern "C" {

const char* eppm_last_error(void) { return g_err.c_str(); }
const char* eppm_version(void) { return "eppm-b200 0.1 (sm_100a)"; }

void eppm_default_params(eppm_params* p) {
    memset(p, 0, sizeof(*p));
    p->pyr_levels = 3;
    p->num_iter = 10;
    p->patch_r = 9;
    p->patch_stride = 2;
    p->search_range = 30;
    p->search_radius_min = 1;
    p->num_rand_guess = 6;
    p->prop_seg_length = 10;
    p->lambda_ad = 0.1f;
    p->lambda_census = 0.3f;
    p->pm_sig_r = 0.1f;
    p->stat_radius = 6;
    p->stat_sim_thresh = 2;
    p->wmf_radius = 4;
    p->wmf_sig_r = 0.02f;
    p->wmf_iters = 20;
    p->blf_sig_s = 5;
    p->blf_sig_r = 0.02f;
    p->rng_mode = EPPM_RNG_XORWOW;
    p->seed = 1234ULL;
}

unsigned long long eppm_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

int eppm_create(eppm_context** out, int device, int h, int w, int max_batch, const eppm_params* params) {
    if (!out || h < 1 || w < 1 || h > 16384 || w > 16384 || max_batch < 1) {
        set_error("eppm_create: bad argument");
        return EPPM_ERR_ARG;
    }
    int ndev = 0;
    if (!cuda_ok(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || device < 0 || device >= ndev) {
        if (g_err.empty()) set_error("eppm_create: no such CUDA device (this library has no CPU path)");
        return EPPM_ERR_CUDA;
    }
    DeviceScope dev_scope(device);
    if (!cuda_ok(cudaGetLastError(), "cudaSetDevice")) return EPPM_ERR_CUDA;
    eppm_context* c = new eppm_context;
    c->device = device;
    c->h = h; c->w = w; c->max_batch = max_batch;
    if (params) c->prm = *params; else eppm_default_params(&c->prm);
    const eppm_params& p = c->prm;
    if (p.pyr_levels < 1 || p.pyr_levels > MAX_LEVELS || p.patch_r != 9 || p.patch_stride < 1 || p.patch_stride > 3 || p.wmf_radius != 4 || p.num_iter < 0 ||
        p.num_rand_guess < 0 || p.num_rand_guess > 16 || p.prop_seg_length < 1 || p.blf_sig_s < 1 || p.blf_sig_s > 10 || p.stat_radius < 0 ||
        p.stat_radius > 16 || (p.rng_mode != EPPM_RNG_XORWOW && p.rng_mode != EPPM_RNG_PHILOX) || p.lambda_ad != 0.1f || p.pm_sig_r != 0.1f) {
        set_error("eppm_create: parameter combination not supported by this build");
        delete c;
        return EPPM_ERR_ARG;
    }
    // ranges the kernels rely on: the search window arithmetic is done in 16-bit (bao_pmflow_kernel.cu:1557-1563), a window must hold at
    // least one pixel, and every sigma / lambda divides
    if (p.search_range < 1 || p.search_range > 32767 || p.search_radius_min < 0 || p.search_radius_min > p.search_range || !(p.lambda_census > 0.f) ||
        !(p.wmf_sig_r > 0.f) || !(p.blf_sig_r > 0.f) || p.wmf_iters < 0 || p.wmf_iters > 1000 || p.stat_sim_thresh < 0 || p.num_iter > 1000 ||
        p.prop_seg_length > 4096 || !(p.lambda_census < 1e6f) || !(p.wmf_sig_r < 1e6f) || !(p.blf_sig_r < 1e6f)) {
        set_error("eppm_create: parameter out of range (search_range 1..32767, 0 <= search_radius_min <= search_range, positive finite lambda_census / "
                  "wmf_sig_r / blf_sig_r, 0 <= wmf_iters, num_iter <= 1000, stat_sim_thresh >= 0)");
        delete c;
        return EPPM_ERR_ARG;
    }
    if (p.subpixel_final && (w % 8 != 0 || p.pyr_levels < 2 || p.patch_stride != 2)) {
        set_error("eppm_create: subpixel_final needs a frame width that is a multiple of 8 (texture pitch), at least two pyramid levels and patch stride 2");
        delete c;
        return EPPM_ERR_ARG;
    }
    // bao_pyr_init_dim (basic/bao_basic.h:196-211): int(double(dim) * pow(double(0.5f), i))
    c->n_levels = p.pyr_levels;
    for (int i = 0; i < c->n_levels; i++) {
        LevelGeom& g = c->lv[i];
        g.h = i == 0 ? h : int(double(h) * pow((double)0.5f, i));
        g.w = i == 0 ? w : int(double(w) * pow((double)0.5f, i));
        g.pw = g.w + 2 * PAD;
        g.ph = g.h + 2 * PAD;
        g.plane = (size_t)g.pw * g.ph;
    }
    const LevelGeom& gc = c->lv[c->n_levels - 1];
    if (gc.w < 1 || gc.h < 1 || (gc.w + p.prop_seg_length - 1) / p.prop_seg_length > 896 || (gc.h + p.prop_seg_length - 1) / p.prop_seg_length > 896) {
        set_error("eppm_create: coarsest level out of range");
        delete c;
        return EPPM_ERR_ARG;
    }
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) c->n_sm = v; }
    c->inplace = p.inplace_filters != 0 || (getenv("EPPM_INPLACE_LEGACY") && atoi(getenv("EPPM_INPLACE_LEGACY")) != 0);
    c->pm_pad_kb = getenv("EPPM_PM_PAD_KB") ? atoi(getenv("EPPM_PM_PAD_KB")) : 0;
    c->profile = getenv("EPPM_PROFILE") && atoi(getenv("EPPM_PROFILE")) != 0;
    {
        // EPPM_STREAM_PRIORITY (measurement knob): CUDA stream priority of this context's compute stream (lower = more urgent), so that
        // two contexts on one device can be ranked when their kernels compete for SMs
        int lo = 0, hi = 0, pr = getenv("EPPM_STREAM_PRIORITY") ? atoi(getenv("EPPM_STREAM_PRIORITY")) : 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        pr = pr < hi ? hi : (pr > lo ? lo : pr);
        if (!cuda_ok(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, pr), "cudaStreamCreate")) { delete c; return EPPM_ERR_CUDA; }
    }
    cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6; i++) cudaEventCreate(&c->ev[i]);
    for (int i = 0; i < 4; i++) cudaEventCreate(&c->ev_k[i]);
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming);
    }

    c->variant = getenv("EPPM_VARIANT") ? atoi(getenv("EPPM_VARIANT")) : 0;
    // ---- arena: two passes (measure, then carve) ----
    const size_t B = max_batch;
    for (int pass = 0; pass < 2; pass++) {
        Arena& A = c->arena;
        A.used = 0;
        if (pass == 1) {
            if (!cuda_ok(cudaMalloc((void**)&A.base, A.size), "cudaMalloc(arena)")) { eppm_destroy(c); return EPPM_ERR_CUDA; }
        } else {
            A.base = nullptr;
        }
        for (int img = 0; img < 2; img++) c->d_rgb[img] = A.take<uint8_t>(B * h * w * 3);
        c->d_flow_out = A.take<float>(B * h * w * 2);
        for (int img = 0; img < 2; img++) c->d_rgb_alt[img] = A.take<uint8_t>(B * h * w * 3);
        c->d_flow_out_alt = A.take<float>(B * h * w * 2);
        for (int i = 0; i < c->n_levels; i++)
            for (int img = 0; img < 2; img++) {
                c->rgba[img][i] = A.take<uchar4>(B * c->lv[i].w * c->lv[i].h);
                c->pix[img][i] = A.take<float4>(B * c->lv[i].plane);
            }
        if (c->n_levels > 2)  // generic blur scratch: the largest source level it can be asked to blur is level 1
            for (int img = 0; img < 2; img++) c->blur_tmp[img] = A.take<uchar4>(B * c->lv[1].w * c->lv[1].h);
        for (int i = 0; i < c->n_levels; i++) c->gauss[i].d_w = A.take<float>(7 * 7 + 1);
        const size_t nc = (size_t)gc.w * gc.h;
        for (int img = 0; img < 2; img++) c->pixT[img] = A.take<float4>(B * gc.plane);
        if (p.patch_stride == 2 && (c->variant & (EPPM_VAR_PROP_Q | EPPM_VAR_PM_Q)))   // parity-split planes (measured alternatives only)
            for (int img = 0; img < 2; img++) c->pixQ[img] = A.take<float4>(B * (size_t)make_qgeom(gc.pw, gc.ph).plane);
        for (int d = 0; d < 2; d++) {
            c->nnf[d] = A.take<short2>(B * nc);
            c->cost[d] = A.take<float>(B * nc);
        }
        c->nnf_tmp = A.take<short2>(B * nc);
        c->occl_list = A.take<int>(B * nc);
        c->occl_count = A.take<int>(64);
        // the PatchMatch-only buffers (random tables, propagation queue, memo) live in a second arena that the first PatchMatch of the
        // context allocates (ensure_pm_buffers): the legacy stage functions create single-level contexts for levels that never run
        // PatchMatch, and rng_search alone is 500 MB at 1920 x 1080
        for (int i = 0; i < c->n_levels; i++) c->flow[i] = A.take<float2>(B * c->lv[i].w * c->lv[i].h);
        c->flow_tmp = A.take<float2>(B * c->lv[0].w * c->lv[0].h);
        if (pass == 0) A.size = A.used;
    }
    c->band_y0 = 0;
    c->band_y1 = gc.h;
    host_luts(c);
    {
        // The smoothing kernel divides by the constant -(sig_r^2).  Its 3-instruction form is used only if it reproduces div.rn
        // for EVERY float the kernel can feed it: dr^2 with dr in {0} U [2^-9, 1] -> [2^-18, 1]; checked once per process and divisor.
        static std::mutex mu;
        static float checked_d = 0.f;
        static int checked_ok = 0;
        std::lock_guard<std::mutex> lk(mu);
        const float d = -(p.blf_sig_r * p.blf_sig_r);
        if (checked_d != d) {
            const float lo = 1.0f / (1 << 20), hi = 2.0f;
            unsigned lob, hib;
            memcpy(&lob, &lo, 4); memcpy(&hib, &hi, 4);
            checked_ok = selftest_const_div(d, lob, hib) == 0;
            checked_d = d;
        }
        c->smooth_fast_div = checked_ok;
    }
    c->vol_ok = p.patch_stride == 2 && build_vol_tab(c->vol_tab);
    for (int l = 0; l + 1 < c->n_levels; l++) c->aff_ok[l] = build_affine_tab(c->aff_tab[l], c->lv[l].pw, c->lv[l].w, c->lv[l].h, p.patch_stride, true);
    for (int l = 0; l + 1 < c->n_levels; l++) {   // spatial weights by sample index (stride-2 sample order: i outer, j inner)
        int s2 = 0;
        for (int i = -PATCH_R; i <= PATCH_R; i += 2)
            for (int j = -PATCH_R; j <= PATCH_R; j += 2, s2++) {
                c->aff_tab[l].gs[s2] = c->cost_lut.gg[i < 0 ? -i : i][j < 0 ? -j : j];
                c->aff_tab[l].boff[0][s2] = ((long long)i * c->lv[l].pw + j) * (long long)sizeof(float4);
                for (int q = 0; q < 3; q++) c->aff_tab[l].boff[q + 1][s2] = (long long)c->aff_tab[l].off[q][s2] * (long long)sizeof(float4);
            }
    }
    build_gauss_tables(c);
    // the random tables are expanded on the first PatchMatch of the context (ensure_rng_tables): the legacy stage functions create
    // contexts for levels that never run PatchMatch
    if (!getenv("EPPM_NO_TMA")) build_smooth_tensor_maps(c);
    {
        // linear textures over the packed planes of the PatchMatch level (both images): element type uint4, point fetch by texel index
        int max_lin = 0;
        cudaDeviceGetAttribute(&max_lin, cudaDevAttrMaxTexture1DLinearWidth, c->device);
        const size_t texels = (size_t)B * gc.plane;
        for (int img = 0; img < 2 && texels <= (size_t)max_lin; img++) {
            const float4* base = c->pix[img][c->n_levels - 1];
            cudaResourceDesc rd = {};
            rd.resType = cudaResourceTypeLinear;
            rd.res.linear.devPtr = const_cast<float4*>(base);
            rd.res.linear.desc = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindUnsigned);
            rd.res.linear.sizeInBytes = texels * sizeof(float4);
            cudaTextureDesc td = {};
            td.readMode = cudaReadModeElementType;
            td.filterMode = cudaFilterModePoint;
            td.addressMode[0] = cudaAddressModeClamp;
            td.normalizedCoords = 0;
            if (cudaCreateTextureObject(&c->tex_pm[img], &rd, &td, nullptr) == cudaSuccess) {
                c->tex_pm_base[img] = base;
                c->tex_pm_texels = texels;
            } else {
                cudaGetLastError();
                c->tex_pm[img] = 0;
            }
        }
    }
    if (!cuda_ok(cudaStreamSynchronize(c->stream), "context setup kernels")) { eppm_destroy(c); return EPPM_ERR_CUDA; }
    *out = c;
    return EPPM_OK;
}

void eppm_destroy(eppm_context* c) {
    if (!c) return;
    DeviceScope dev_scope(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->tile_comm) eppm_tiled_shutdown(c);
    for (int img = 0; img < 2; img++)
        if (c->tex_pm[img]) cudaDestroyTextureObject(c->tex_pm[img]);
    if (c->arena.base) cudaFree(c->arena.base);
    for (int i = 0; i < 2; i++)
        if (c->subpix_census[i]) cudaFree(c->subpix_census[i]);
    if (c->subpix_nnf) cudaFree(c->subpix_nnf);
    if (c->pm_arena.base) cudaFree(c->pm_arena.base);
    for (int i = 0; i < 2; i++)
        if (c->h_pinned_in[i]) cudaFreeHost(c->h_pinned_in[i]);
    if (c->h_pinned_out) cudaFreeHost(c->h_pinned_out);
    for (int i = 0; i < 6; i++)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 4; i++)
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
    for (int i = 0; i < 2; i++) {
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
        if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

int eppm_num_levels(const eppm_context* c) { return c ? c->n_levels : EPPM_ERR_ARG; }
int eppm_level_dims(const eppm_context* c, int level, int* h, int* w) {
    if (!c || level < 0 || level >= c->n_levels) return EPPM_ERR_ARG;
    if (h) *h = c->lv[level].h;
    if (w) *w = c->lv[level].w;
    return EPPM_OK;
}
void* eppm_stream(eppm_context* c) { return c ? (void*)c->stream : nullptr; }
int eppm_synchronize(eppm_context* c) {
    if (!c) return EPPM_ERR_ARG;
    DeviceScope dev_scope(c->device);
    return cuda_ok(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize") ? EPPM_OK : EPPM_ERR_CUDA;
}

static int check_batch(eppm_context* c, const void* a, const void* b, int n) {
    if (!c || !a || !b || n < 1 || n > c->max_batch) {
        set_error("bad argument (null pointer or batch size out of range)");
        return EPPM_ERR_ARG;
    }
    return EPPM_OK;
}

int eppm_stage_prepare(eppm_context* c, const uint8_t* d_img1, const uint8_t* d_img2, int n) {
    int rc = check_batch(c, d_img1, d_img2, n);
    if (rc) return rc;
    DeviceScope dev_scope(c->device);
    c->n_cur = n;
    run_prepare(c, d_img1, d_img2, n);
    return cuda_ok(cudaGetLastError(), "prepare") ? EPPM_OK : EPPM_ERR_CUDA;
}
int eppm_stage_patchmatch(eppm_context* c) {
    if (!c || c->n_cur < 1) { set_error("patchmatch before prepare"); return EPPM_ERR_STATE; }
    DeviceScope dev_scope(c->device);
    run_patchmatch(c);
    return cuda_ok(cudaGetLastError(), "patchmatch") ? EPPM_OK : EPPM_ERR_CUDA;
}
int eppm_stage_patchmatch_partial(eppm_context* c, int n_steps) {
    if (!c || c->n_cur < 1) { set_error("patchmatch before prepare"); return EPPM_ERR_STATE; }
    DeviceScope dev_scope(c->device);
    run_patchmatch_dirs(c, 2, n_steps);
    return cuda_ok(cudaGetLastError(), "patchmatch") ? EPPM_OK : EPPM_ERR_CUDA;
}
int eppm_stage_consistency(eppm_context* c) {
    if (!c || c->n_cur < 1) { set_error("consistency before prepare"); return EPPM_ERR_STATE; }
    DeviceScope dev_scope(c->device);
    run_consistency(c);
    return cuda_ok(cudaGetLastError(), "consistency") ? EPPM_OK : EPPM_ERR_CUDA;
}
int eppm_stage_c2f(eppm_context* c, float* d_flow) {
    if (!c || c->n_cur < 1) { set_error("c2f before prepare"); return EPPM_ERR_STATE; }
    DeviceScope dev_scope(c->device);
    run_c2f(c, d_flow);
    return cuda_ok(cudaGetLastError(), "c2f") ? EPPM_OK : EPPM_ERR_CUDA;
}

int eppm_compute_batch_device(eppm_context* c, const uint8_t* d_img1, const uint8_t* d_img2, int n, float* d_flow) {
    int rc = check_batch(c, d_img1, d_img2, n);
    if (rc) return rc;
    if (!d_flow) { set_error("null output"); return EPPM_ERR_ARG; }
    DeviceScope dev_scope(c->device);
    c->n_cur = n;
    cudaStream_t s = c->stream;
    if (c->profile) cudaEventRecord(c->ev[0], s);
    run_prepare(c, d_img1, d_img2, n);
    if (c->profile) cudaEventRecord(c->ev[1], s);
    run_patchmatch(c);
    if (c->profile) cudaEventRecord(c->ev[2], s);
    run_consistency(c);
    if (c->profile) cudaEventRecord(c->ev[3], s);
    run_c2f(c, d_flow);
    if (c->profile) cudaEventRecord(c->ev[4], s);
    return cuda_ok(cudaGetLastError(), "compute_batch") ? EPPM_OK : EPPM_ERR_CUDA;
}

// Video stream: n_pairs consecutive pairs (frame f, frame f+1) of a list of n_pairs+1 frames; every frame's
// pyramid / census / packed planes are built once and shared by the two pairs it belongs to.
int eppm_compute_stream_device(eppm_context* c, const uint8_t* d_frames, int n_pairs, float* d_flow) {
    if (!c || !d_frames || !d_flow || n_pairs < 1 || n_pairs + 1 > c->max_batch) {
        set_error("bad argument (null pointer, or n_pairs + 1 frames exceed max_batch)");
        return EPPM_ERR_ARG;
    }
    DeviceScope dev_scope(c->device);
    run_prepare_frames(c, d_frames, n_pairs + 1);
    // image 2 of pair f is frame f+1: alias the image-2 plane bases one plane behind the image-1 bases for this call
    float4* keep_pix[MAX_LEVELS];
    float4* keep_pixT = c->pixT[1];
    float4* keep_pixQ = c->pixQ[1];
    for (int l = 0; l < c->n_levels; l++) {
        keep_pix[l] = c->pix[1][l];
        c->pix[1][l] = c->pix[0][l] + c->lv[l].plane;
    }
    c->pixT[1] = c->pixT[0] + c->lv[c->n_levels - 1].plane;
    c->pixQ[1] = c->pixQ[0] + make_qgeom(c->lv[c->n_levels - 1].pw, c->lv[c->n_levels - 1].ph).plane;
    c->n_cur = n_pairs;
    run_patchmatch(c);
    run_consistency(c);
    run_c2f(c, d_flow);
    for (int l = 0; l < c->n_levels; l++) c->pix[1][l] = keep_pix[l];
    c->pixT[1] = keep_pixT;
    c->pixQ[1] = keep_pixQ;
    return cuda_ok(cudaGetLastError(), "compute_stream") ? EPPM_OK : EPPM_ERR_CUDA;
}

int eppm_last_stage_ms(eppm_context* c, float out[5]) {
    if (!c || !c->profile) return EPPM_ERR_STATE;
    DeviceScope dev_scope(c->device);
    if (!cuda_ok(cudaEventSynchronize(c->ev[4]), "event sync")) return EPPM_ERR_CUDA;
    for (int i = 0; i < 4; i++) cudaEventElapsedTime(&out[i], c->ev[i], c->ev[i + 1]);
    cudaEventElapsedTime(&out[4], c->ev[0], c->ev[4]);
    return EPPM_OK;
}

int eppm_last_kernel_ms(eppm_context* c, int which, float* ms) {
    if (!c || !c->profile || !ms || which < 0 || which > 1) return EPPM_ERR_STATE;
    DeviceScope dev_scope(c->device);
    if (!cuda_ok(cudaEventSynchronize(c->ev_k[2 * which + 1]), "event sync")) return EPPM_ERR_CUDA;
    return cuda_ok(cudaEventElapsedTime(ms, c->ev_k[2 * which], c->ev_k[2 * which + 1]), "event elapsed") ? EPPM_OK : EPPM_ERR_CUDA;
}

// Host-buffer entry point.  n may exceed max_batch: the batch is processed in chunks of max_batch pairs, double buffered --
// while chunk k computes on the context's stream, chunk k+1 is uploaded and chunk k-1's flow is downloaded on the copy stream
// (true overlap needs pinned host memory; pageable memory still works, the copies then serialise on the host).
int eppm_compute_batch_host(eppm_context* c, const uint8_t* img1, const uint8_t* img2, int n, float* flow) {
    if (!c || !img1 || !img2 || n < 1) { set_error("bad argument (null pointer or empty batch)"); return EPPM_ERR_ARG; }
    if (!flow) { set_error("null output"); return EPPM_ERR_ARG; }
    DeviceScope dev_scope(c->device);
    const size_t px = (size_t)c->h * c->w;
    cudaStream_t s = c->stream, cs = c->copy_stream;
    uint8_t* in[2][2] = {{c->d_rgb[0], c->d_rgb[1]}, {c->d_rgb_alt[0], c->d_rgb_alt[1]}};
    float* out[2] = {c->d_flow_out, c->d_flow_out_alt};
    const int B = c->max_batch, n_chunks = (n + B - 1) / B;
    auto upload = [&](int k) {
        const int off = k * B, m = n - off < B ? n - off : B, slot = k & 1;
        if (k >= 2) cudaStreamWaitEvent(cs, c->ev_done[slot], 0);  // the chunk that used this slot has consumed its inputs
        cudaMemcpyAsync(in[slot][0], img1 + (size_t)off * px * 3, (size_t)m * px * 3, cudaMemcpyHostToDevice, cs);
        cudaMemcpyAsync(in[slot][1], img2 + (size_t)off * px * 3, (size_t)m * px * 3, cudaMemcpyHostToDevice, cs);
        cudaEventRecord(c->ev_h2d[slot], cs);
    };
    upload(0);
    for (int k = 0; k < n_chunks; k++) {
        const int off = k * B, m = n - off < B ? n - off : B, slot = k & 1;
        cudaStreamWaitEvent(s, c->ev_h2d[slot], 0);
        if (k >= 2) cudaStreamWaitEvent(s, c->ev_d2h[slot], 0);  // the previous result in this slot has left the device
        int rc = eppm_compute_batch_device(c, in[slot][0], in[slot][1], m, out[slot]);
        if (rc) {   // copies of earlier chunks into the caller's memory may still be in flight: drain both streams before giving up
            const std::string keep = g_err;
            cudaStreamSynchronize(cs);
            cudaStreamSynchronize(s);
            set_error(keep);
            return rc;
        }
        cudaEventRecord(c->ev_done[slot], s);
        if (k + 1 < n_chunks) upload(k + 1);
        cudaStreamWaitEvent(cs, c->ev_done[slot], 0);
        cudaMemcpyAsync(flow + (size_t)off * px * 2, out[slot], (size_t)m * px * 2 * sizeof(float), cudaMemcpyDeviceToHost, cs);
        cudaEventRecord(c->ev_d2h[slot], cs);
    }
    if (!cuda_ok(cudaStreamSynchronize(cs), "compute_batch_host (copy stream)")) return EPPM_ERR_CUDA;
    return cuda_ok(cudaStreamSynchronize(s), "compute_batch_host") ? EPPM_OK : EPPM_ERR_CUDA;
}

// Video stream with HOST buffers: frames [n_frames][h][w][3] u8, flow [n_frames - 1][h][w][2] f32.  The stream is cut into
// chunks of max_batch - 1 pairs (max_batch frames; neighbouring chunks share one frame, uploaded with both); while chunk k computes,
// chunk k+1 uploads and chunk k-1's flow downloads on the copy stream.  Each frame's pyramid / census / packed planes are built once per
// chunk.  Replaces a loop of set_data + compute_flow over consecutive pairs (main.cpp:59-65 in a video loop).
int eppm_compute_stream_host(eppm_context* c, const uint8_t* frames, int n_frames, float* flow) {
    if (!c || !frames || !flow || n_frames < 2) { set_error("bad argument (null pointer or fewer than two frames)"); return EPPM_ERR_ARG; }
    if (c->max_batch < 2) { set_error("eppm_compute_stream_host needs a context with max_batch >= 2"); return EPPM_ERR_ARG; }
    DeviceScope dev_scope(c->device);
    const size_t px = (size_t)c->h * c->w;
    cudaStream_t s = c->stream, cs = c->copy_stream;
    uint8_t* in[2] = {c->d_rgb[0], c->d_rgb_alt[0]};
    float* out[2] = {c->d_flow_out, c->d_flow_out_alt};
    const int P = c->max_batch - 1, n_pairs = n_frames - 1, n_chunks = (n_pairs + P - 1) / P;
    auto upload = [&](int k) {
        const int f0 = k * P, m = n_pairs - f0 < P ? n_pairs - f0 : P, slot = k & 1;
        if (k >= 2) cudaStreamWaitEvent(cs, c->ev_done[slot], 0);
        cudaMemcpyAsync(in[slot], frames + (size_t)f0 * px * 3, (size_t)(m + 1) * px * 3, cudaMemcpyHostToDevice, cs);
        cudaEventRecord(c->ev_h2d[slot], cs);
    };
    upload(0);
    for (int k = 0; k < n_chunks; k++) {
        const int f0 = k * P, m = n_pairs - f0 < P ? n_pairs - f0 : P, slot = k & 1;
        cudaStreamWaitEvent(s, c->ev_h2d[slot], 0);
        if (k >= 2) cudaStreamWaitEvent(s, c->ev_d2h[slot], 0);
        int rc = eppm_compute_stream_device(c, in[slot], m, out[slot]);
        if (rc) {
            const std::string keep = g_err;
            cudaStreamSynchronize(cs);
            cudaStreamSynchronize(s);
            set_error(keep);
            return rc;
        }
        cudaEventRecord(c->ev_done[slot], s);
        if (k + 1 < n_chunks) upload(k + 1);
        cudaStreamWaitEvent(cs, c->ev_done[slot], 0);
        cudaMemcpyAsync(flow + (size_t)f0 * px * 2, out[slot], (size_t)m * px * 2 * sizeof(float), cudaMemcpyDeviceToHost, cs);
        cudaEventRecord(c->ev_d2h[slot], cs);
    }
    if (!cuda_ok(cudaStreamSynchronize(cs), "compute_stream_host (copy stream)")) return EPPM_ERR_CUDA;
    return cuda_ok(cudaStreamSynchronize(s), "compute_stream_host") ? EPPM_OK : EPPM_ERR_CUDA;
}

long eppm_read_plane(eppm_context* c, int which, int level, int pair, void* host_out) {
    if (!c || !host_out || level < 0 || level >= c->n_levels || pair < 0 || pair >= c->max_batch) return EPPM_ERR_ARG;
    DeviceScope dev_scope(c->device);
    if (!cuda_ok(cudaStreamSynchronize(c->stream), "read_plane sync")) return EPPM_ERR_CUDA;
    const LevelGeom& g = c->lv[level];
    const LevelGeom& gc = c->lv[c->n_levels - 1];
    const size_t n = (size_t)g.w * g.h, nc = (size_t)gc.w * gc.h;
    cudaError_t e = cudaSuccess;
    long bytes = 0;
    switch (which) {
    case EPPM_PLANE_RGBA1: case EPPM_PLANE_RGBA2:
        bytes = n * 4;
        e = cudaMemcpy(host_out, c->rgba[which - EPPM_PLANE_RGBA1][level] + pair * n, bytes, cudaMemcpyDeviceToHost);
        break;
    case EPPM_PLANE_CENSUS1: case EPPM_PLANE_CENSUS2: {
        unsigned char* tmp = nullptr;
        bytes = n;
        if (!cuda_ok(cudaMalloc((void**)&tmp, n), "cudaMalloc")) return EPPM_ERR_CUDA;
        k_extract_census<<<dim3((g.w + 127) / 128, g.h), 128, 0, c->stream>>>(c->pix[which - EPPM_PLANE_CENSUS1][level] + pair * g.plane, g.pw, tmp, g.w, g.h);
        cudaStreamSynchronize(c->stream);
        e = cudaMemcpy(host_out, tmp, n, cudaMemcpyDeviceToHost);
        cudaFree(tmp);
        break;
    }
    case EPPM_PLANE_NNF_FWD: case EPPM_PLANE_NNF_BWD:
        bytes = nc * 4;
        e = cudaMemcpy(host_out, c->nnf[which - EPPM_PLANE_NNF_FWD] + pair * nc, bytes, cudaMemcpyDeviceToHost);
        break;
    case EPPM_PLANE_COST_FWD: case EPPM_PLANE_COST_BWD:
        bytes = nc * 4;
        e = cudaMemcpy(host_out, c->cost[which - EPPM_PLANE_COST_FWD] + pair * nc, bytes, cudaMemcpyDeviceToHost);
        break;
    case EPPM_PLANE_FLOW:
        bytes = n * 8;
        e = cudaMemcpy(host_out, c->flow[level] + pair * n, bytes, cudaMemcpyDeviceToHost);
        break;
    case EPPM_PLANE_FLOW_TMP:  // scratch plane, laid out with the dims of `level` by the step that wrote it
        bytes = n * 8;
        e = cudaMemcpy(host_out, c->flow_tmp + pair * n, bytes, cudaMemcpyDeviceToHost);
        break;
    default:
        return EPPM_ERR_ARG;
    }
    return cuda_ok(e, "read_plane copy") ? bytes : EPPM_ERR_CUDA;
}

// ---- spatial tiling of one large frame across GPUs: every rank holds the full pyramids and fields, owns a band of
// coarsest-level rows aligned to the propagation segment length, and exchanges one boundary row per column pass (eppm_b200/tiled.py).
int eppm_set_band(eppm_context* c, int band, int n_bands) {
    if (!c || n_bands < 1 || band < 0 || band >= n_bands) { set_error("eppm_set_band: bad argument"); return EPPM_ERR_ARG; }
    const LevelGeom& gc = c->lv[c->n_levels - 1];
    const int sl = c->prm.prop_seg_length;
    const int n_seg = (gc.h + sl - 1) / sl;
    if (n_bands > 1 && n_seg / n_bands < 2) { set_error("eppm_set_band: fewer than two segments per band"); return EPPM_ERR_ARG; }
    const int base = n_seg / n_bands, extra = n_seg % n_bands;
    const int s0 = band * base + (band < extra ? band : extra), s1 = s0 + base + (band < extra ? 1 : 0);
    c->band_y0 = s0 * sl;
    c->band_y1 = s1 * sl < gc.h ? s1 * sl : gc.h;
    return EPPM_OK;
}
int eppm_band_rows(eppm_context* c, int level, int* y0, int* y1) {
    if (!c || level < 0 || level >= c->n_levels || !y0 || !y1) return EPPM_ERR_ARG;
    band_rows(c, level, y0, y1);
    return EPPM_OK;
}
int eppm_tiled_pm_steps(eppm_context* c, int first_step, int end_step) {
    if (!c || c->n_cur < 1) { set_error("patchmatch before prepare"); return EPPM_ERR_STATE; }
    DeviceScope dev_scope(c->device);
    run_patchmatch_dirs(c, 2, end_step, first_step);
    return cuda_ok(cudaGetLastError(), "tiled patchmatch") ? EPPM_OK : EPPM_ERR_CUDA;
}
int eppm_tiled_c2f_step(eppm_context* c, int level, int kind) {
    if (!c || c->n_cur < 1 || level < 0 || level >= c->n_levels || kind < 0 || kind > 2 || (kind == 0 && level >= c->n_levels - 1)) {
        set_error("eppm_tiled_c2f_step: bad argument");
        return EPPM_ERR_ARG;
    }
    DeviceScope dev_scope(c->device);
    run_c2f_step(c, level, kind, nullptr);
    return cuda_ok(cudaGetLastError(), "tiled c2f") ? EPPM_OK : EPPM_ERR_CUDA;
}
void* eppm_device_plane(eppm_context* c, int which, int level) {
    if (!c || level < 0 || level >= c->n_levels) return nullptr;
    switch (which) {
    case EPPM_PLANE_NNF_FWD: case EPPM_PLANE_NNF_BWD: return c->nnf[which - EPPM_PLANE_NNF_FWD];
    case EPPM_PLANE_COST_FWD: case EPPM_PLANE_COST_BWD: return c->cost[which - EPPM_PLANE_COST_FWD];
    case EPPM_PLANE_FLOW: return c->flow[level];
    case EPPM_PLANE_FLOW_TMP: return c->flow_tmp;
    }
    return nullptr;
}

int eppm_selftest_affine_sites(int w, int h, int pw, int* table_out) {
    AffineTab t;
    if (w < 1 || h < 1 || pw < 1) return EPPM_ERR_ARG;
    const bool ok = build_affine_tab(t, pw, w, h);
    if (ok && table_out)
        for (int q = 0; q < 3; q++) memcpy(table_out + q * 100, t.off[q], 100 * sizeof(int));   // stride 2: [3][100]
    return ok ? 1 : 0;
}
int eppm_selftest_affine_sites_stride(int w, int h, int pw, int stride, int* table_out) {
    AffineTab t;
    if (w < 1 || h < 1 || pw < 1 || stride < 1 || stride > 3) return EPPM_ERR_ARG;
    const bool ok = build_affine_tab(t, pw, w, h, stride);
    const int n = (2 * PATCH_R) / stride + 1;
    if (ok && table_out)
        for (int q = 0; q < 3; q++) memcpy(table_out + q * n * n, t.off[q], (size_t)n * n * sizeof(int));
    return ok ? 1 : 0;
}
int eppm_selftest_volume_tables(int* box_out, int* t_out, unsigned* used_out) {
    static VolTab v;
    const bool ok = build_vol_tab(v);
    if (ok && box_out)
        for (int r = 0; r < 10; r++) { box_out[4 * r] = v.xlo[r]; box_out[4 * r + 1] = v.ylo[r]; box_out[4 * r + 2] = v.bx[r]; box_out[4 * r + 3] = v.by[r]; }
    if (ok && t_out) memcpy(t_out, v.T, sizeof(v.T));
    if (ok && used_out) memcpy(used_out, v.used, sizeof(v.used));
    return ok ? 1 : 0;
}
int eppm_refine_uses_site_table(eppm_context* c, int level) {
    if (!c || level < 0 || level >= c->n_levels) return EPPM_ERR_ARG;
    return c->aff_ok[level] && !(c->variant & EPPM_VAR_REFINE_GENERIC);
}

long long eppm_selftest_const_div(float d, unsigned lo_bits, unsigned hi_bits) { return selftest_const_div(d, lo_bits, hi_bits); }
int eppm_smooth_uses_fast_div(eppm_context* c) { return c ? c->smooth_fast_div : EPPM_ERR_ARG; }
int eppm_smooth_uses_tma(eppm_context* c) { return c ? c->tmap_ok[0] : EPPM_ERR_ARG; }

long eppm_write_plane(eppm_context* c, int which, int level, int pair, const void* host_in) {
    if (!c || !host_in || level < 0 || level >= c->n_levels || pair < 0 || pair >= c->max_batch) return EPPM_ERR_ARG;
    DeviceScope dev_scope(c->device);
    if (!cuda_ok(cudaStreamSynchronize(c->stream), "write_plane sync")) return EPPM_ERR_CUDA;
    const LevelGeom& g = c->lv[level];
    const LevelGeom& gc = c->lv[c->n_levels - 1];
    const size_t n = (size_t)g.w * g.h, nc = (size_t)gc.w * gc.h;
    cudaError_t e;
    long bytes;
    switch (which) {
    case EPPM_PLANE_NNF_FWD: case EPPM_PLANE_NNF_BWD:
        bytes = nc * 4;
        e = cudaMemcpy(c->nnf[which - EPPM_PLANE_NNF_FWD] + pair * nc, host_in, bytes, cudaMemcpyHostToDevice);
        break;
    case EPPM_PLANE_COST_FWD: case EPPM_PLANE_COST_BWD:
        bytes = nc * 4;
        e = cudaMemcpy(c->cost[which - EPPM_PLANE_COST_FWD] + pair * nc, host_in, bytes, cudaMemcpyHostToDevice);
        break;
    case EPPM_PLANE_FLOW:
        bytes = n * 8;
        e = cudaMemcpy(c->flow[level] + pair * n, host_in, bytes, cudaMemcpyHostToDevice);
        break;
    default:
        return EPPM_ERR_ARG;
    }
    // a caller-written field voids what the propagation remembers about candidates it has already scored
    if (which != EPPM_PLANE_FLOW && c->prop_memo) cudaMemset(c->prop_memo, 0xff, (size_t)2 * c->max_batch * nc * sizeof(int4));
    return cuda_ok(e, "write_plane copy") ? bytes : EPPM_ERR_CUDA;
}

}  // extern "C"
