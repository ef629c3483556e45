// Host-side context of libeppm_b200 (not part of the public ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/eppm.h"
#include "eppm_device.cuh"

namespace eppm {

constexpr int MAX_LEVELS = 6;

// A/B switches read once from the environment variable EPPM_VARIANT at eppm_create (measurement knobs; every setting computes the same bits)
enum {
    EPPM_VAR_REFINE_GENERIC = 1,   // plane-fitting refine: coordinates computed per sample instead of the verified site table
    EPPM_VAR_REFINE_NOGROUP = 2,   // table kernel with the per-sample __expf fix-up instead of the grouped test
    EPPM_VAR_SEARCH_SERIAL = 4,    // random search: one guess at a time instead of all guesses side by side
    EPPM_VAR_SMOOTH_2ROW = 16,     // smoothing: the generic two-rows-per-thread kernel instead of four rows in packed pairs
    EPPM_VAR_PROP_NOCOMPACT = 32,  // propagation: skip per thread, but no compaction of the remaining evaluations across the CTA
    EPPM_VAR_PROP_CTA = 64,        // propagation: CTA-local lock-step kernels (with compaction) instead of the global work queue
    EPPM_VAR_REFINE_9WARP3 = 256,  // table refine with one candidate (four models) per thread: nine warps per 32 pixels, 72 registers, 3 CTAs per SM
    EPPM_VAR_SEARCH_TEX3 = 512,    // random search: the three wide-window guesses gather their target side through the texture unit
    EPPM_VAR_SEARCH_NOTEX = 1024,  //   ... none (default: the two widest)
    EPPM_VAR_SEARCH_SPLIT3 = 65536,  //   ... three passes of two (64 registers, 8 CTAs)
    EPPM_VAR_REFINE_PK = 2048,         // table refine with the four models of a candidate in packed pairs (measured slower: 8.82 vs 8.29 ms per 1080p pair at level 0; packed FP32x2 holds the FMA pipe two cycles)
    EPPM_VAR_REFINE_PK_BRANCH = 4096,  // packed-pair refine that branches around candidate rows outside the image instead of scoring them at a clamped centre
    EPPM_VAR_PROP_THREAD = 8192,       // propagation queue scored by one THREAD per evaluation (round-1 default) instead of one warp per evaluation
    EPPM_VAR_PROP_CHAIN = 16384,       // propagation as barrier-free segment chains, 32 chains per warp (measured slower than the queue: 5.70 vs 5.26 ms per pair)
    EPPM_VAR_SEARCH_WARP = 32768,      // random search with one WARP per evaluation (measured slower: neighbouring pixels' narrow guesses already coalesce in the thread-per-pixel kernel)
    EPPM_VAR_PROP_NOMEMO = 262144,     // propagation: score candidates the pixel has scored before (the reference does; the memo skips them, same outcome)
    EPPM_VAR_PM_Q = 524288,            // PatchMatch kernels read parity-split (Q) planes with 256-bit loads, two samples per request (measured slower: 4.79-4.88 vs 4.67 ms per pair;
                                       //   a 256-bit request costs the L1 as many wavefronts as two 128-bit ones)
    EPPM_VAR_REFINE_COLUMN = 1048576,  // table refine with warp = candidate column, thread = three candidate rows (round-1 default) instead of warp = candidate row
    EPPM_VAR_PROP_Q = 2097152,         // propagation: the warp-per-evaluation scoring kernel reads the parity-split planes (dense sample rows; measured slower: 4.45 vs 4.24 ms per pair)
    EPPM_VAR_PROP_WARP_FULL = 131072,  // propagation queue scored by one warp per evaluation with ALL samples staged in shared memory (13 KB per warp starves the L1)
    EPPM_VAR_REFINE_VOLUME = 8388608,  // refine with the AD + census terms of a CTA computed once per (image-1 column, displacement) into shared memory (k_c2f_refine_vol; measured slower: 8.75 vs 7.54 ms per pair at level 0, bound by L1 wavefronts)
    EPPM_VAR_CENSUS_NOTMA = 16777216,  // census + pack: nine clamped loads and nine luminances per thread (round 1) instead of a TMA-staged tile whose luminances are formed once
    EPPM_VAR_REFINE_NOFASTW = 4194304, // refine: every patch row with the __expf fix-up test (default: only the first row, then a test-free loop where that provably changes no bit)
    EPPM_VAR_PROP_NOSKIP = 8,      // propagation: evaluate candidates that equal the current target (the reference does)
};

struct LevelGeom {
    int w, h;     // bao_pyr_init_dim: int(double(dim) * pow(0.5, level))  (basic/bao_basic.h:196-211)
    int pw, ph;   // padded plane dims: w + 2*PAD, h + 2*PAD
    size_t plane; // pw*ph
};

struct GaussTab {   // tap weights of one pyramid blur, computed on the device with the reference's expression
    float* d_w;     // [(2r+1)^2]
    int r;
};

struct Arena {
    char* base = nullptr;
    size_t size = 0, used = 0;
    template <class T>
    T* take(size_t n) {
        size_t bytes = (n * sizeof(T) + 511) & ~size_t(511);   // 512 B: the texture alignment (linear textures are bound over arena planes)
        T* p = reinterpret_cast<T*>(base + used);
        used += bytes;
        return p;
    }
};

struct AffineTab {
    int off[3][361];   // dy * pw + dx of affine model 1..3 at sample (i, j) of the patch (100 samples at stride 2, 361 at 1, 49 at 3), relative to the candidate centre; i outer, j inner
    // at most one site whose x offset is NOT a function of (i, j, model) alone (stride 1: model 3 at (-7, -2) lies 3e-8 from an integer):
    // the table holds its offset without dx and the kernel computes dx = floor(fma(i, exc_ci, fma(j, exc_cj, float(X)))) - X per thread
    int exc_s;         // sample index of that site, -1 = none
    int exc_q;         // its model (0..2)
    float exc_cj, exc_ci;
    long long boff[4][100];   // stride 2: BYTE offset of the site of model 0..3 (0 = identity) at sample s: the 64-bit add takes it straight from the constant bank
    float gs[100];     // stride 2: spatial weight G[|j|] * G[|i|] of sample s (CostLut::gg by sample index: one constant load, no index arithmetic)
};

// Shared AD + census volume of the refine kernel (k_c2f_refine_vol, stride 2).  The AD + census half of a patch sample depends only on the
// image-1 pixel q1 = x + j of patch row i and on the displacement (dX, dY) = D(x) + candidate + site offset of the model -- not on the pixel x
// that asks for it.  Per patch row r = (i + 9) / 2 the displacements all models and candidates of a CTA can ask for (relative to the CTA's
// smallest integer flow, flows spreading by at most 1 in x and y inside the CTA) fill a box of bx[r] x by[r] "lines"; a line holds the term
// for the CTA's 50 image-1 columns.  T[q][s] is the byte offset of (model q, sample s) inside the volume for the centre candidate (m = 1,
// n = 1) of a lane whose flow is the CTA minimum, column of lane 0; used[r][sx + 2 sy] marks the lines some lane can read when the flows
// spread by sx / sy.
constexpr unsigned CEN_TILE_W = 40, CEN_TILE_H = 10;   // census kernel: 32 x 8 pixels + halo = 34 x 10, fetched from 3 pixels further left: a TMA box must start on a 16-byte boundary (probed: tools/probe_tma_u32.cu)
constexpr int VOL_COLS = 32 + 2 * PATCH_R;   // image-1 columns of a CTA of 32 pixels
constexpr int VOL_MAX_LINES = 104;           // >= max over rows of bx * by (98 for the reference's coefficient sets; checked by build_vol_tab)
struct VolTab {
    int T[4][100];
    int xlo[10], ylo[10], bx[10], by[10];
    unsigned used[10][4][4];
};

struct SmoothLut {
    float g[21];     // expf(-i^2 / sig_s^2), i = 0..2*sig_s  (bao_pmflow_refine_kernel.cu:809-813)
    float pad_[3];
};
struct WmfLut {
    float g[8];      // expf(-i^2 / (WMF_SIG_S^2)), i = 0..WMF_RADIUS (:270-275)
};

}  // namespace eppm

struct eppm_context {
    int device = 0;
    int h = 0, w = 0, max_batch = 0;
    int n_cur = 0;                 // pairs of the batch currently resident
    eppm_params prm;
    int n_levels = 0;
    eppm::LevelGeom lv[eppm::MAX_LEVELS];
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    bool profile = false;
    cudaEvent_t ev[6] = {};
    cudaEvent_t ev_k[4] = {};      // [0],[1] bracket k_c2f_refine at level 0; [2],[3] k_flow_smooth of the final pass
    float stage_ms[5] = {};

    eppm::Arena arena;
    eppm::Arena pm_arena;                        // PatchMatch-only buffers, allocated by the first PatchMatch of the context (ensure_pm_buffers)
    // inputs staged on the device for the host-buffer API
    uint8_t* d_rgb[2] = {nullptr, nullptr};      // [B][h][w][3]
    float* d_flow_out = nullptr;                 // [B][h][w][2]
    uint8_t* d_rgb_alt[2] = {nullptr, nullptr};  // second staging set: chunk k+1 uploads while chunk k computes
    float* d_flow_out_alt = nullptr;
    cudaEvent_t ev_h2d[2] = {}, ev_done[2] = {}, ev_d2h[2] = {};
    uint8_t* h_pinned_in[2] = {nullptr, nullptr};
    float* h_pinned_out = nullptr;
    // pyramid
    uchar4* rgba[2][eppm::MAX_LEVELS] = {};      // [B][h_l][w_l] dense
    uchar4* blur_tmp[2] = {nullptr, nullptr};    // scratch for pyramid levels beyond the 2-octave fast path
    float4* pix[2][eppm::MAX_LEVELS] = {};       // [B][ph_l][pw_l] packed float rgb + census
    float4* pixT[2] = {nullptr, nullptr};        // column-major copies of the coarsest level ([B][pw][ph]) for row propagation
    float4* pixQ[2] = {nullptr, nullptr};        // parity-split copies of the coarsest level (QGeom, eppm_device.cuh): 256-bit sample-pair loads
    cudaTextureObject_t tex_pm[2] = {0, 0};      // linear uint4 textures over pix[img][coarsest]: scattered gathers of the random search
    const float4* tex_pm_base[2] = {nullptr, nullptr};
    size_t tex_pm_texels = 0;
    eppm::GaussTab gauss[eppm::MAX_LEVELS];      // [0] pre-blur, [i] blur feeding level i
    // PatchMatch state at the coarsest level: index = dir (0 fwd, 1 bwd)
    short2* nnf[2] = {nullptr, nullptr};         // [B][h_c][w_c]
    float* cost[2] = {nullptr, nullptr};
    short2* nnf_tmp = nullptr;                   // snapshot buffer for the in-place filters
    int* occl_list = nullptr;                    // compacted occluded pixels for WMF: [B*h_c*w_c]
    int* occl_count = nullptr;                   // [2] device counters
    short2* prop_prev = nullptr;                 // propagation work queue: running target per (pair, direction, line, segment)
    int4* prop_queue = nullptr;                  //   evaluations of the current lock-step
    int4* prop_memo = nullptr;                   //   last candidate scored per (pair, direction, pixel) and pass direction: [2*B][h_c][w_c]
    int* prop_count = nullptr;                   //   queue length per (pass, step): [num_iter * 4 * seg_len]
    int n_sm = 148;
    short2* rng_init = nullptr;                  // [h_c][w_c] initial targets (same for every pair/direction)
    short2* rng_search = nullptr;                // [num_iter][num_guess][h_c][w_c] raw (short(r1), short(r2))
    // flow pyramid (+ snapshot buffer)
    float2* flow[eppm::MAX_LEVELS] = {};
    float2* flow_tmp = nullptr;                  // level-0 sized
    eppm::CostLut cost_lut;
    eppm::SmoothLut smooth_lut;
    eppm::WmfLut wmf_lut;
    CUtensorMap tmap_pix0[eppm::MAX_LEVELS];     // TMA descriptors of the image-1 packed planes (smoothing tile loads)
    int tmap_ok[eppm::MAX_LEVELS] = {};
    CUtensorMap tmap_refine[eppm::MAX_LEVELS];   // ... and of the refine kernel's image-1 tile (default refine: EPPM_REFINE_MODE=18)
    int tmap_refine_ok[eppm::MAX_LEVELS] = {};
    CUtensorMap tmap_rgba[2][eppm::MAX_LEVELS];  // ... and of the dense RGBA levels (tile + halo of the census kernel)
    int tmap_rgba_ok[2][eppm::MAX_LEVELS] = {};
    int tmap_box_h = 0;                          // tile height the tensor maps were encoded for
    void* tile_comm = nullptr;                   // ncclComm_t of the tiling group (tiled.cu); rank / size below
    int tile_rank = 0, tile_world = 0;
    int band_y0 = 0, band_y1 = 0;                // rows of the coarsest level this context owns (whole level unless tiled across GPUs)
    eppm::VolTab vol_tab;                        // shared AD + census volume tables of the refine kernel (stride 2 only)
    int vol_ok = 0;
    eppm::AffineTab aff_tab[eppm::MAX_LEVELS];   // per level (pitch): verified sample-site tables of the plane-fitting refine
    int aff_ok[eppm::MAX_LEVELS] = {};
    int variant = 0;                             // EPPM_VARIANT bit mask (A/B switches for measurements, see EPPM_VAR_*)
    int smooth_fast_div = 0;                     // set at create time when the constant-division fast path was verified exact
    int pm_pad_kb = 0;                           // EPPM_PM_PAD_KB: dynamic shared memory (KB) the PatchMatch scoring kernels reserve without using it = residency cap (co-scheduling)
    unsigned char* subpix_census[2] = {nullptr, nullptr};   // [2h][2w] census planes of the bicubic-upsampled images (subpixel_final; allocated on first use)
    short2* subpix_nnf = nullptr;                           // [h][w] integer targets of one pair
    int inplace = 0;                             // eppm_params::inplace_filters or EPPM_INPLACE_LEGACY=1: the three racy filters of the reference run in place (legacy_inplace.cu)
    int rng_ready = 0;                           // rng_init / rng_search expanded (lazily, before the first PatchMatch)
};

namespace eppm {
void set_error(const std::string& s);
bool cuda_ok(cudaError_t e, const char* what);
extern std::atomic<unsigned long long> g_launches;
#define EPPM_LAUNCH_COUNT(n) (eppm::g_launches += (n))

// stage drivers (each enqueues on ctx->stream)
void run_prepare(eppm_context* c, const uint8_t* d_img1, const uint8_t* d_img2, int n);
void run_patchmatch(eppm_context* c);
void run_patchmatch_dirs(eppm_context* c, int n_dirs, int n_steps = 1 << 30, int first_step = 0);
bool run_patchmatch_scaled(eppm_context* c, float* d_scale);   // baoCudaPatchMatch_Scaled: forward direction of pair 0; d_scale = dense [h][w] device plane
bool run_patchmatch_planefitting(eppm_context* c);   // baoCudaPatchMatch_PlaneFitting: forward direction of pair 0, coarsest-level planes
void run_c2f_step(eppm_context* c, int level, int kind, float2* out);
bool build_smooth_tensor_maps(eppm_context* c);
void op_subpix_final(eppm_context* c);   // eppm_params::subpixel_final: the reference's sub-pixel stage on flow_tmp (level 0), pair by pair (subpix.cu)
void band_rows(const eppm_context* c, int level, int* y0, int* y1);
void run_consistency(eppm_context* c);
void run_c2f(eppm_context* c, float* d_flow_out);
void build_rng_tables(eppm_context* c);
bool ensure_pm_buffers(eppm_context* c);   // second arena: random tables, propagation queue / memo (lazily, legacy single-level contexts rarely need them)
void ensure_rng_tables(eppm_context* c);   // builds them on the context's stream the first time a PatchMatch is queued
void build_gauss_tables(eppm_context* c);
bool build_affine_tab(AffineTab& t, int pw, int w, int h, int stride = 2, bool allow_exception = false);
bool build_vol_tab(VolTab& v);   // stride 2; false if the boxes do not fit VOL_MAX_LINES

// building blocks reused by the legacy stage ABI (foreign buffers)
void op_lr_check(cudaStream_t s, short2* nnf, float* cost, const short2* nnf2, int w, int h, int n);
void op_outlier_removal(cudaStream_t s, const short2* src, short2* dst, float* cost, int w, int h, int n, int R, int sim);
void wmf_sweeps(eppm_context* c, short2*& cur, short2*& other, const float4* pix, size_t plane, int pw, int w, int h, int n, int iters,
                bool only_occlusion);
void op_fill_holes(cudaStream_t s, const short2* src, short2* dst, const float4* pix, size_t plane, int pw, int w, int h, int n);
void op_nnf_to_flow(cudaStream_t s, const short2* nnf, float2* flow, int w, int h, int n);
void op_image_bilateral(eppm_context* c, uchar4* dst, const uchar4* src, size_t pitch_bytes, const float4* pix1, const eppm::LevelGeom& g);
void op_flow_bilateral_upsample(eppm_context* c, float2* dst, const float4* pix1, const eppm::LevelGeom& g, const float2* small, int ws, float ratio);
void op_refine(eppm_context* c, const float4* pix1, const float4* pix2, const LevelGeom& g, const float2* coarse, int ws, int hs, int upsample,
               float2* out, int n, int y0 = 0, int y1 = -1);
long long selftest_const_div(float d, unsigned lo_bits, unsigned hi_bits);
void op_smooth(eppm_context* c, const float2* src, float2* dst, const float4* pix1, const LevelGeom& g, int n, int y0 = 0, int y1 = -1);
// in-place forms of the reference's racy filters (legacy_inplace.cu; opt-in)
void op_outlier_inplace(eppm_context* c, short2* nnf, float* cost, int w, int h, int n);
void op_wmf_inplace(eppm_context* c, short2* nnf, const float4* pix, size_t plane, int pw, int w, int h, int n, int iters, bool only_occlusion);
void op_smooth_inplace(eppm_context* c, float2* flow, const float4* pix1, const LevelGeom& g, int n);
void op_preblur_rgba(eppm_context* c, const uchar4* src1, const uchar4* src2, size_t pitch_bytes);
void op_pyramid_and_pack(eppm_context* c, int n, int two = 2);
void run_prepare_frames(eppm_context* c, const uint8_t* d_frames, int n_frames);
void op_pack_foreign(cudaStream_t s, const uchar4* rgba, size_t rgba_pitch_bytes, const unsigned char* census, size_t census_pitch_bytes, float4* pix,
                     const LevelGeom& g);
void op_transpose_plane(cudaStream_t s, const float4* src, float4* dst, const LevelGeom& g, int n_img);
void op_split_plane(cudaStream_t s, const float4* src, float4* dst, const LevelGeom& g, int n_img);
void op_extract_census(cudaStream_t s, const float4* pix, const LevelGeom& g, unsigned char* out, size_t out_pitch_bytes);
void k_pack_planes(cudaStream_t s, const uchar4* rgba, size_t rgba_pitch_bytes, size_t rgba_img_stride_bytes, float4* pix,
                   const LevelGeom& g, int n_img);
}  // namespace eppm
