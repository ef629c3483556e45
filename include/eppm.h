/* eppm.h — C ABI of libeppm_b200.so: the B200-native (sm_100a) implementation of EPPM's
 * dense-correspondence hot path (census -> pyramid -> coarse-level PatchMatch -> consistency pass ->
 * coarse-to-fine plane-fitting refine + joint-bilateral smoothing).
 *
 * Two layers are exported, both `extern "C"`, plain pointers and sizes only:
 *
 *  (1) the batched, stream-ordered context API `eppm_*` declared below — what a new caller binds
 *      (ctypes / cgo / JNI stubs are shown in INTEGRATION.md);
 *  (2) the reference's own stage functions `baoCuda*` with the reference's signatures and buffer
 *      layouts (declared in include/eppm_legacy_abi.h), so that the reference's host class
 *      (bao_flow_patchmatch_multiscale_cuda.cpp:40-62 of linchaobao/EPPM) links against this
 *      library unchanged.
 *
 * There is no CPU fallback: every entry point fails with EPPM_ERR_CUDA when no sm_100 device is present.
 * All citations `file:line` refer to the upstream reference linchaobao/EPPM.
 */
#ifndef EPPM_H_
#define EPPM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPPM_OK 0
#define EPPM_ERR_ARG (-1)     /* bad argument (null pointer, size out of range, batch > max_batch) */
#define EPPM_ERR_CUDA (-2)    /* CUDA runtime error; text via eppm_last_error() */
#define EPPM_ERR_STATE (-3)   /* call sequence error (e.g. compute before set) */

#define EPPM_RNG_XORWOW 0 /* the reference's stream: curand XORWOW, seed 1234, one subsequence per 16x16
                             block (bao_pmflow_kernel.cu:50-109,1519-1586) -> bit-exact NNF */
#define EPPM_RNG_PHILOX 1 /* counter-based Philox4x32-10 keyed by (seed, iteration, guess, pixel) */

/* Algorithm parameters. The reference fixes all of them as macros (defs.h:31-76, and
 * PROP_SEG_LENGTH bao_pmflow_kernel.cu:979, DIFF_THRESH/STAT_SIM_THRESH/POSTPROC_BLF_SIG_R
 * bao_pmflow_refine_kernel.cu:51,147,752); eppm_default_params() returns exactly those values. */
typedef struct eppm_params {
    int pyr_levels;          /* PYR_MAX_DEPTH 3 */
    int num_iter;            /* NUM_ITER 10 */
    int patch_r;             /* PATCH_R 9        (kernels are specialised for 9) */
    int patch_stride;        /* 2  (pixel skipping, bao_pmflow_kernel.cu:269,272); 1..3 supported */
    int search_range;        /* SEARCH_RANGE 30 */
    int search_radius_min;   /* SEARCH_RADIUS_MIN 1 */
    int num_rand_guess;      /* NUM_RAND_GUESS 6 */
    int prop_seg_length;     /* PROP_SEG_LENGTH 10 */
    float lambda_ad;         /* LAMBDA_AD 0.1 */
    float lambda_census;     /* LAMBDA_CENSUS 0.3 */
    float pm_sig_r;          /* PM_SIG_R 0.1 */
    int stat_radius;         /* STAT_RADIUS 6 */
    int stat_sim_thresh;     /* STAT_SIM_THRESH 2 */
    int wmf_radius;          /* WMF_RADIUS 4 */
    float wmf_sig_r;         /* WMF_SIG_R 0.02 */
    int wmf_iters;           /* 20 (bao_flow_patchmatch_multiscale_cuda.cpp:239) */
    int blf_sig_s;           /* POSTPROC_BLF_SIG_S 5 (radius = 2*sig_s) */
    float blf_sig_r;         /* POSTPROC_BLF_SIG_R 0.02 */
    int rng_mode;            /* EPPM_RNG_XORWOW */
    unsigned long long seed; /* 1234 (bao_pmflow_kernel.cu:68) */
    int inplace_filters;     /* 0 (default): outlier removal, weighted median and flow smoothing read a snapshot and write a second
                                buffer (deterministic).  1: they update IN PLACE with the reference's 16x16 launch geometry and
                                per-thread raster walk (bao_pmflow_refine_kernel.cu:149-193, 206-286, 764-826), reproducing the
                                reference's read-while-write race as far as the hardware schedules it the same way; slower, and
                                like the reference not guaranteed reproducible across GPUs.  EPPM_INPLACE_LEGACY=1 in the
                                environment forces 1 for every context. */
    int subpixel_final;      /* 0 (default).  1: the reference's optional sub-pixel stage (baoCudaCensusTransform_Bicubic + baoCudaSubpixRefine,
                                bao_pmflow_refine_kernel.cu:395-634 -- declared by its host class, never placed in its pipeline) runs on the
                                integer flow of the level-0 refine, in front of the two level-0 smoothing passes, at full resolution as the
                                reference's buffers are sized (…cuda.cpp:135-136).  Needs w % 8 == 0 (texture pitch) and >= 2 pyramid levels;
                                costs about three times the rest of the pipeline.  Not combined with spatial tiling. */
    int reserved[6];
} eppm_params;

typedef struct eppm_context eppm_context; /* opaque; one per (device, h, w, params) */

void eppm_default_params(eppm_params* p);

/* Create a context on `device` for frames of h x w (RGB u8) and up to `max_batch` pairs per call.
 * Replaces bao_flow_patchmatch_multiscale_cuda::init(h,w) (…cuda.cpp:112-157): all device memory
 * (one arena), the RNG tables and LUTs are set up here, nothing is allocated per pair. */
int eppm_create(eppm_context** out, int device, int h, int w, int max_batch, const eppm_params* params);
void eppm_destroy(eppm_context* ctx);
const char* eppm_last_error(void);
const char* eppm_version(void);

/* Level geometry, identical to bao_pyr_init_dim (basic/bao_basic.h:196-211). */
int eppm_num_levels(const eppm_context* ctx);
int eppm_level_dims(const eppm_context* ctx, int level, int* h, int* w);

/* One batch, HOST buffers (the call a user of the reference's class makes, batched):
 *   img1,img2 : [n][h][w][3] u8 interleaved RGB (bao_alloc layout, main.cpp:42-57)
 *   flow      : [n][h][w][2] f32 interleaved (u,v) — what compute_flow returns as u[][] / v[][]
 * n may exceed max_batch: the batch is then processed in chunks of max_batch pairs with the upload of chunk k+1 and the
 * download of chunk k-1 overlapped with the compute of chunk k (pinned host memory recommended).
 * H2D, all stages, D2H; returns after the result is in `flow`.
 * Replaces set_data + compute_flow (…cuda.cpp:159-168,217-306). */
int eppm_compute_batch_host(eppm_context* ctx, const uint8_t* img1, const uint8_t* img2, int n, float* flow);

/* Same with DEVICE-resident inputs/outputs (same layouts), stream-ordered on the context's stream;
 * returns without synchronising. */
int eppm_compute_batch_device(eppm_context* ctx, const uint8_t* d_img1, const uint8_t* d_img2, int n, float* d_flow);
/* Video stream: d_frames = [n_pairs + 1][h][w][3] consecutive frames (device), d_flow = [n_pairs][h][w][2]; pair f is
 * (frame f -> frame f+1).  Each frame's pyramid, census and packed planes are built once and used by both pairs it belongs to
 * (the reference rebuilds them per pair, bao_flow_patchmatch_multiscale_cuda.cpp:159-168).  n_pairs + 1 <= max_batch. */
int eppm_compute_stream_device(eppm_context* ctx, const uint8_t* d_frames, int n_pairs, float* d_flow);
/* The same with HOST buffers: frames [n_frames][h][w][3] u8, flow [n_frames - 1][h][w][2] f32; n_frames may exceed max_batch (the stream
 * is processed in chunks of max_batch - 1 pairs, uploads / downloads of neighbouring chunks overlapped with compute; pinned host memory
 * recommended).  max_batch >= 2.  Replaces a video loop of set_data + compute_flow (main.cpp:59-65). */
int eppm_compute_stream_host(eppm_context* ctx, const uint8_t* frames, int n_frames, float* flow);
int eppm_synchronize(eppm_context* ctx);
void* eppm_stream(eppm_context* ctx); /* cudaStream_t */

/* Staged execution for stage-level parity tests and profiling (device-resident):
 *   eppm_stage_prepare     : RGB -> pre-blur, pyramid, census, packed planes      (…refine_kernel.cu:1060-1071)
 *   eppm_stage_patchmatch  : both directions at the coarsest level                (…pmflow_kernel.cu:1760-1826)
 *   eppm_stage_consistency : LR check, outlier removal, WMF, hole filling, NNF->flow (…cuda.cpp:233-258)
 *   eppm_stage_c2f         : for each finer level: upsample x2, plane-fitting refine, smoothing; final smoothing
 * Call in this order after eppm_stage_prepare. */
int eppm_stage_prepare(eppm_context* ctx, const uint8_t* d_img1, const uint8_t* d_img2, int n);
int eppm_stage_patchmatch(eppm_context* ctx);
int eppm_stage_consistency(eppm_context* ctx);
int eppm_stage_c2f(eppm_context* ctx, float* d_flow);

/* Read-outs of intermediate planes of pair `pair` (dense host copies; synchronises):
 *   EPPM_PLANE_RGBA1/2   uchar4 [h_l][w_l]      pyramid level image
 *   EPPM_PLANE_CENSUS1/2 u8     [h_l][w_l]
 *   EPPM_PLANE_NNF_FWD/BWD short2 [h_c][w_c]    absolute targets at the coarsest level
 *   EPPM_PLANE_COST_FWD/BWD f32 [h_c][w_c]
 *   EPPM_PLANE_FLOW      float2 [h_l][w_l]
 * Returns bytes written or a negative error. */
enum {
    EPPM_PLANE_RGBA1 = 0, EPPM_PLANE_RGBA2 = 1, EPPM_PLANE_CENSUS1 = 2, EPPM_PLANE_CENSUS2 = 3,
    EPPM_PLANE_NNF_FWD = 4, EPPM_PLANE_NNF_BWD = 5, EPPM_PLANE_COST_FWD = 6, EPPM_PLANE_COST_BWD = 7,
    EPPM_PLANE_FLOW = 8,
    EPPM_PLANE_FLOW_TMP = 9 /* scratch flow plane (level-0 sized): output of a tiled refine step / of the final smoothing */
};
long eppm_read_plane(eppm_context* ctx, int which, int level, int pair, void* host_out);

/* Test hooks: overwrite an NNF / COST / FLOW plane of pair `pair` from a dense host array (same shapes as
 * eppm_read_plane), and run only the first `n_steps` launch groups of the PatchMatch stage
 * (1 = random field + initial cost, then per iteration 4 propagation passes and 1 random search). */
long eppm_write_plane(eppm_context* ctx, int which, int level, int pair, const void* host_in);
int eppm_stage_patchmatch_partial(eppm_context* ctx, int n_steps);

/* Spatial tiling of ONE large frame pair across GPUs (one context per GPU, all holding the full frame; batch of 1).
 *   eppm_set_band(ctx, band, n_bands): this context owns band `band` of `n_bands` row bands of the coarsest level, aligned to the
 *       propagation segment length so column segments never straddle a band.  (band 0 of 1 = whole frame, the default.)
 *   eppm_band_rows: the rows [y0, y1) of `level` the band maps to.
 *   eppm_tiled_pm_steps(ctx, first, end): PatchMatch launch groups [first, end) on the band (numbering of
 *       eppm_stage_patchmatch_partial: 0 = random field + cost, 1+5*it+k = pass k of iteration it, k = 4 random search).
 *       Between a row pass and the following column pass the caller copies ONE boundary row of both NNF planes from the
 *       neighbouring band (row y0-1 before a forward column pass, row y1 before a reverse one) -- the halo exchange.
 *   eppm_tiled_c2f_step(ctx, level, kind): kind 0 = upsample+refine (-> FLOW_TMP), 1 = smoothing FLOW_TMP -> FLOW[level],
 *       2 = final smoothing FLOW[0] -> FLOW_TMP, each on the band's rows; the caller all-gathers the written rows in between.
 *   eppm_device_plane: device address of pair 0 of a plane, for the caller's NCCL / peer copies.
 * eppm_b200/tiled.py drives this with torch.distributed; results are bit-identical to the untiled run. */
int eppm_set_band(eppm_context* ctx, int band, int n_bands);
int eppm_band_rows(eppm_context* ctx, int level, int* y0, int* y1);
int eppm_tiled_pm_steps(eppm_context* ctx, int first_step, int end_step);
int eppm_tiled_c2f_step(eppm_context* ctx, int level, int kind);
void* eppm_device_plane(eppm_context* ctx, int which, int level);

/* The same tiling driven BY THE LIBRARY (BASELINE config 4): one process and one context (max_batch >= 1) per GPU, every rank calls the same
 * functions in the same order.  Halo rows and band gathers travel through NCCL (NVLink / NVSwitch), enqueued on the context's stream; NCCL is
 * loaded at run time (libnccl.so.2; EPPM_NCCL_LIB overrides), the library does not link against it.
 *   eppm_tiled_unique_id(id): 128 bytes (ncclUniqueId) from ONE rank, which the caller hands to all ranks by any means (MPI, torch.distributed, a file)
 *   eppm_tiled_init(ctx, rank, world, id): collective; joins the communicator and fixes the rank's band (whole propagation segments; needs at least two per band)
 *   eppm_compute_tiled_device / _host: collective; inputs are the SAME full frame pair on every rank, the full flow [h][w][2] arrives on every rank
 *       (host variant: flow may be NULL on ranks that do not want the copy).  Bit-identical to eppm_compute_batch_* on one GPU.
 *   eppm_tiled_shutdown(ctx): leaves the communicator (also done by eppm_destroy). */
int eppm_tiled_unique_id(void* id_out_128_bytes);
int eppm_tiled_init(eppm_context* ctx, int rank, int world, const void* unique_id);
int eppm_compute_tiled_device(eppm_context* ctx, const uint8_t* d_img1, const uint8_t* d_img2, float* d_flow);
int eppm_compute_tiled_host(eppm_context* ctx, const uint8_t* img1, const uint8_t* img2, float* flow);
int eppm_tiled_shutdown(eppm_context* ctx);

/* Exhaustive device self-test: number of floats x with bit patterns in [lo_bits, hi_bits) for which the 3-instruction
 * constant division (q0 = x*r; q = fma(fma(q0,-d,x), r, q0), r = RN(1/d)) differs from div.rn(x, d); 0 = exact everywhere.
 * eppm_smooth_uses_fast_div reports whether a context's smoothing kernel was allowed to use it. */
long long eppm_selftest_const_div(float d, unsigned lo_bits, unsigned hi_bits);
int eppm_smooth_uses_fast_div(eppm_context* ctx);
/* 1 when the smoothing kernel stages its colour tile with TMA (cp.async.bulk.tensor); EPPM_NO_TMA=1 in the environment disables it. */
int eppm_smooth_uses_tma(eppm_context* ctx);

/* The plane-fitting refine (d_bilateral_refine_flow_planefitting, bao_pmflow_kernel.cu:2005-2069) samples the three affine patch
 * models at floor(fma(i, Cy, fma(j, Cx, float(X)))).  Those sites are tabulated per (i, j, model) relative to the
 * candidate centre; the table is used only after eppm_create has checked, with the host's correctly rounded fmaf, that it
 * reproduces the expression for EVERY coordinate X the level can produce.  eppm_selftest_affine_sites runs that check on the host
 * (no GPU needed) for a level of w x h with plane pitch pw: returns 1 and writes the 3 x 100 element offsets (dy * pw + dx) to
 * table_out (may be NULL) when the table is exact, 0 when it is not (the refine then computes the coordinates per sample).
 * eppm_refine_uses_site_table reports what a context's refine kernel does at `level`. */
int eppm_selftest_affine_sites(int w, int h, int pw, int* table_out);
/* The same for patch stride 1, 2 or 3: n = 19, 10 or 7 samples per patch row, table_out = 3 x n*n offsets.  Strides 2 and 3 verify at
 * every size; stride 1 does not as a pure table (returns 0): one site, model 3 at (i, j) = (-7, -2), lies 3e-8 from an integer -- a
 * context at stride 1 tabulates the other 1082 sites and computes that one per thread with the reference's arithmetic. */
int eppm_selftest_affine_sites_stride(int w, int h, int pw, int stride, int* table_out);
int eppm_refine_uses_site_table(eppm_context* ctx, int level);
/* Tables of the refine variant with a shared AD + census volume (EPPM_VARIANT bit 8388608): per patch row r = 0..9 the box of
 * displacement "lines" (box_out: 10 x {xlo, ylo, bx, by}), the byte offset of (model q, sample s) inside the volume for the centre candidate
 * of a lane at the CTA's minimum flow (t_out: 4 x 100) and the lines some lane can read when the CTA's flows spread by sx / sy
 * (used_out: 10 x 4 x 4 words, index [r][sx + 2 sy]).  Host only (no CUDA call); returns 1 when the boxes fit the kernel's volume. */
int eppm_selftest_volume_tables(int* box_out, int* t_out, unsigned* used_out);

/* Evaluation against ground truth on the device (d_flow, d_gt: [n][h][w][2] f32 device arrays; out: n host records; synchronises).
 *   epe / aae_deg : bao_calc_flow_error (basic/bao_flow_tools.cpp:64-111): mean end-point error and mean angular error (degrees) over the
 *                   pixels inside `border` whose ground truth is known (|.| <= 1e9) and non-zero; n_valid = their number
 *   outlier_frac  : bao_calc_flow_error_percentage (:114-141): share of the pixels with known ground truth (n_known) whose end-point error
 *                   exceeds outlier_thresh
 * Per-pixel arithmetic is the reference's; the sums are double and order-deterministic. */
typedef struct eppm_flow_error {
    double epe, aae_deg, outlier_frac;
    long long n_valid, n_known;
} eppm_flow_error;
int eppm_eval_flow(eppm_context* ctx, const float* d_flow, const float* d_gt, int n, int border, float outlier_thresh, eppm_flow_error* out);
/* Middlebury .flo files (3rdparty/middlebury/flowIO.cpp:56-160) from / to the interleaved (u, v) host layout of this API: "PIEH", int32 w,
 * int32 h, h*w*2 float32.  eppm_read_flo with flow_uv == NULL only returns the dimensions. */
int eppm_write_flo(const char* path, const float* flow_uv, int h, int w);
int eppm_read_flo(const char* path, float* flow_uv, int* h, int* w, size_t capacity_floats);

/* Number of kernel launches issued by this library since the counter was last reset (bench.py's gpu_launches). */
unsigned long long eppm_launch_count(int reset);

/* Per-stage device time of the last eppm_compute_batch_* call in ms (prepare, patchmatch, consistency, c2f, total);
 * valid only when the context was created with EPPM_PROFILE=1 in the environment. */
int eppm_last_stage_ms(eppm_context* ctx, float out[5]);
/* Device time of one kernel of the last call, measured with CUDA events on the stream it was launched on
 * (EPPM_PROFILE=1): which = 0 plane-fitting refine at level 0 (the dominant kernel), 1 = final flow smoothing. */
int eppm_last_kernel_ms(eppm_context* ctx, int which, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* EPPM_H_ */
