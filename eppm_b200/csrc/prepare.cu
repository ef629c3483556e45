// Stage "prepare": RGB u8 -> 5x5 pre-blur (= pyramid level 0) -> Gaussian pyramid -> 3x3 census ->
// packed float planes.  Restates baoCudaPatchMatchMultiscalePrepare (bao_pmflow_refine_kernel.cu:1060-1071),
// _d_bao_gauss_filter<uchar4> (basic/bao_basic_cuda.cuh:437-467), _d_bao_bilinear_resize<uchar4> (:565-601),
// bao_cuda_construct_gauss_pyramid_pitched (:642-664) and d_census_transform3x3 (bao_pmflow_census_kernel.cu:39-90).
//
// Pyramid schedule AS COMPILED: `int n = log(0.25)/log(ratio)` (basic/bao_basic_cuda.cuh:648) divides a double by the
// FLOAT logarithm of 0.5f (the .cu picks the float overload), giving 1.99999999 -> n = 1, not 2.  So level 1 is the
// sigma=1 / radius=3 blur of level 0 kept at (2x+1,2y+1), and every further level i is the sigma=1 / radius=3 blur of
// level i-1 resized by (float)pow(.5,i)*W0/W[i-1] (:658-661) -- exactly 0.5 whenever the widths halve exactly.
//
// B200 design: the reference blurs the whole source level and then keeps one pixel in 4 (the "bilinear" resize at ratio
// 0.5 has weight exactly 1 on texel (2x+1,2y+1)); here the 7x7 blur is evaluated only at those sites (identical bytes,
// 4x fewer taps).  Non-halving sizes take the generic blur + true bilinear path.  The tap weights are
// per-offset constants; they are produced once per context by a device kernel that evaluates the reference's own
// expression (__expf(-(float)(dy*dy+dx*dx)/sigma)), so they are the same bits the reference recomputes per pixel.
#include "eppm_internal.h"

namespace eppm {

// ---------------------------------------------------------------------------------------------------------
// Gaussian tap tables
__global__ void k_gauss_table(float* out, float sigma, int r) {
    // basic/bao_basic_cuda.cuh:444,452: sigma = sigma*sigma*2; weight = __expf(-(float)(dy*dy+dx*dx)/sigma)
    float s = sigma * sigma;
    s = s + s;
    const int n = 2 * r + 1;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
        int dy = i / n - r, dx = i % n - r;
        out[i] = __expf(-(float)(dy * dy + dx * dx) / s);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // `sum += weight` in tap order (:459) gives the same value at every pixel: fold it into the table
        float sum = 0.f;
        for (int i = 0; i < n * n; i++) sum = __fadd_rn(sum, out[i]);
        out[n * n] = sum;
    }
}

void build_gauss_tables(eppm_context* c) {
    // level 0: pre-blur sigma .5 radius 2 (bao_pmflow_refine_kernel.cu:1063); every other level: sigma 1, radius 3
    // (basic/bao_basic_cuda.cuh:647-661 with baseSigma = 1/ratio-1 = 1 and n = 1, see the header comment).
    for (int i = 0; i < c->n_levels; i++) {
        float sigma = i == 0 ? 0.5f : 1.0f;
        int r = i == 0 ? 2 : 3;
        c->gauss[i].r = r;
        k_gauss_table<<<1, 256, 0, c->stream>>>(c->gauss[i].d_w, sigma, r);
        EPPM_LAUNCH_COUNT(1);
    }
}

// ---------------------------------------------------------------------------------------------------------
// 5x5 pre-blur fused with RGB -> RGBA (bao_rgb2rgba, basic/bao_basic_cuda.h:258-267: alpha = 0).
// One thread per output pixel; the 20x20 source tile (+2 halo) is staged in shared memory as uchar4.
// SRC_RGBA = false: packed RGB u8 [B][h][w][3] (the eppm_* API); true: uchar4 rows `pitch` bytes apart (legacy stage ABI).
constexpr int PB_T = 16;
template <bool SRC_RGBA>
__global__ void __launch_bounds__(PB_T* PB_T) k_preblur_rgb(const uint8_t* __restrict__ rgb0, const uint8_t* __restrict__ rgb1,
                                                           uchar4* __restrict__ out0, uchar4* __restrict__ out1, int w, int h,
                                                           const float* __restrict__ wtab, size_t pitch, int two) {
    __shared__ uchar4 tile[PB_T + 4][PB_T + 4];
    __shared__ float sw[26];
    // two = 2: grid.z = pair*2 + image (two source arrays); two = 1: grid.z = frame of ONE flat frame list (video stream)
    const int img = two == 2 ? (blockIdx.z & 1) : 0, b = two == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    const uint8_t* src = (img ? rgb1 : rgb0) + (SRC_RGBA ? (size_t)b * h * pitch : (size_t)b * h * w * 3);
    uchar4* dst = (img ? out1 : out0) + (size_t)b * h * w;
    const int tid = threadIdx.y * PB_T + threadIdx.x;
    if (tid < 26) sw[tid] = wtab[tid];
    const int x0 = blockIdx.x * PB_T - 2, y0 = blockIdx.y * PB_T - 2;
    for (int i = tid; i < (PB_T + 4) * (PB_T + 4); i += PB_T * PB_T) {
        int ty = i / (PB_T + 4), tx = i % (PB_T + 4);
        int cy = max(0, min(h - 1, y0 + ty)), cx = max(0, min(w - 1, x0 + tx));  // :449-450 clamp
        if (SRC_RGBA) {
            tile[ty][tx] = *reinterpret_cast<const uchar4*>(src + (size_t)cy * pitch + (size_t)cx * 4);
        } else {
            const uint8_t* p = src + ((size_t)cy * w + cx) * 3;
            tile[ty][tx] = make_uchar4(p[0], p[1], p[2], 0);
        }
    }
    __syncthreads();
    const int x = blockIdx.x * PB_T + threadIdx.x, y = blockIdx.y * PB_T + threadIdx.y;
    if (x >= w || y >= h) return;
    float vx = 0.f, vy = 0.f, vz = 0.f;
#pragma unroll
    for (int dy = 0; dy < 5; dy++)
#pragma unroll
        for (int dx = 0; dx < 5; dx++) {
            const uchar4 p = tile[threadIdx.y + dy][threadIdx.x + dx];
            const float wgt = sw[dy * 5 + dx];
            vx = __fmaf_rn(wgt, (float)p.x, vx);  // :454-457  val += pix * weight  (FFMA weight, float(u8), val)
            vy = __fmaf_rn(wgt, (float)p.y, vy);
            vz = __fmaf_rn(wgt, (float)p.z, vz);
        }
    const float sum = sw[25];
    // :460-466  val /= sum; uchar = trunc(val)
    dst[(size_t)y * w + x] = make_uchar4((unsigned char)__float2uint_rz(__fdiv_rn(vx, sum)), (unsigned char)__float2uint_rz(__fdiv_rn(vy, sum)),
                                         (unsigned char)__float2uint_rz(__fdiv_rn(vz, sum)), 0);
}

// ---------------------------------------------------------------------------------------------------------
// Pyramid level i from its source level at ratio exactly 0.5: out(x,y) = blur(src)(2*(x+1)-1, 2*(y+1)-1), i.e. the one
// texel the reference's resize keeps (basic/bao_basic_cuda.cuh:565-601 with integral fx, dx = dy = 0).  Blur = :437-467.
// One thread per OUTPUT pixel, taps read straight from the source level (L1/L2-cache resident).
__global__ void __launch_bounds__(128) k_pyr_decimate(const uchar4* __restrict__ l0a, const uchar4* __restrict__ l0b, uchar4* __restrict__ outa,
                                                      uchar4* __restrict__ outb, int w0, int h0, int w, int h, int step, int r,
                                                      const float* __restrict__ wtab, int two) {
    extern __shared__ float sw[];
    const int n = 2 * r + 1;
    for (int i = threadIdx.x; i <= n * n; i += blockDim.x) sw[i] = wtab[i];
    __syncthreads();
    const int img = two == 2 ? (blockIdx.z & 1) : 0, b = two == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    const uchar4* src = (img ? l0b : l0a) + (size_t)b * h0 * w0;
    uchar4* dst = (img ? outb : outa) + (size_t)b * h * w;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    // :572-577  fx = (x+1)*(1/ratio) - 1 (integral), xx = int(fx), u = clamp(xx)
    const int sx = max(0, min(w0 - 1, (x + 1) * step - 1)), sy = max(0, min(h0 - 1, (y + 1) * step - 1));
    float vx = 0.f, vy = 0.f, vz = 0.f;
    for (int dy = -r; dy <= r; dy++) {
        const uchar4* row = src + (size_t)max(0, min(h0 - 1, sy + dy)) * w0;
        const float* wr = sw + (dy + r) * n + r;
        for (int dx = -r; dx <= r; dx++) {
            const uchar4 p = __ldg(row + max(0, min(w0 - 1, sx + dx)));
            const float wgt = wr[dx];
            vx = __fmaf_rn(wgt, (float)p.x, vx);
            vy = __fmaf_rn(wgt, (float)p.y, vy);
            vz = __fmaf_rn(wgt, (float)p.z, vz);
        }
    }
    const float sum = sw[n * n];
    dst[(size_t)y * w + x] = make_uchar4((unsigned char)__float2uint_rz(__fdiv_rn(vx, sum)), (unsigned char)__float2uint_rz(__fdiv_rn(vy, sum)),
                                         (unsigned char)__float2uint_rz(__fdiv_rn(vz, sum)), 0);
}

// Generic path for levels whose resize ratio is not exactly 0.5 (odd source dimensions, basic/bao_basic_cuda.cuh:658-661):
// full-resolution blur of the source level, then the true bilinear resize.
__global__ void k_blur_full(const uchar4* __restrict__ src0, uchar4* __restrict__ dst0, int w, int h, int r, const float* __restrict__ wtab) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const uchar4* src = src0 + (size_t)blockIdx.z * w * h;
    uchar4* dst = dst0 + (size_t)blockIdx.z * w * h;
    const int n = 2 * r + 1;
    float vx = 0.f, vy = 0.f, vz = 0.f;
    for (int dy = -r; dy <= r; dy++)
        for (int dx = -r; dx <= r; dx++) {
            const uchar4 p = src[(size_t)max(0, min(h - 1, y + dy)) * w + max(0, min(w - 1, x + dx))];
            const float wgt = wtab[(dy + r) * n + dx + r];
            vx = __fmaf_rn(wgt, (float)p.x, vx);
            vy = __fmaf_rn(wgt, (float)p.y, vy);
            vz = __fmaf_rn(wgt, (float)p.z, vz);
        }
    const float sum = wtab[n * n];
    dst[(size_t)y * w + x] = make_uchar4((unsigned char)__float2uint_rz(__fdiv_rn(vx, sum)), (unsigned char)__float2uint_rz(__fdiv_rn(vy, sum)),
                                         (unsigned char)__float2uint_rz(__fdiv_rn(vz, sum)), 0);
}

__global__ void k_resize_u8(uchar4* __restrict__ dst0, int ow, int oh, const uchar4* __restrict__ src0, int w, int h, float ratio) {
    // basic/bao_basic_cuda.cuh:565-601
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= ow) return;
    const uchar4* src = src0 + (size_t)blockIdx.z * w * h;
    uchar4* dst = dst0 + (size_t)blockIdx.z * ow * oh;
    const float div_scale = 1.f / ratio;
    const float fx = __fmaf_rn((float)(x + 1), div_scale, -1.f), fy = __fmaf_rn((float)(y + 1), div_scale, -1.f);
    const int xx = (int)fx, yy = (int)fy;
    const float dx = fmaxf(fminf(__fsub_rn(fx, (float)xx), 1.f), 0.f), dy = fmaxf(fminf(__fsub_rn(fy, (float)yy), 1.f), 0.f);
    float rx = 0.f, ry = 0.f, rz = 0.f;
#pragma unroll
    for (int m = 0; m <= 1; m++)
#pragma unroll
        for (int n = 0; n <= 1; n++) {
            const int u = max(0, min(w - 1, xx + m)), v = max(0, min(h - 1, yy + n));
            const float s = __fmul_rn(fabsf(__fsub_rn((float)(1 - m), dx)), fabsf(__fsub_rn((float)(1 - n), dy)));
            const uchar4 p = src[(size_t)v * w + u];
            rx = __fmaf_rn((float)p.x, s, rx);
            ry = __fmaf_rn((float)p.y, s, ry);
            rz = __fmaf_rn((float)p.z, s, rz);
        }
    dst[(size_t)y * ow + x] = make_uchar4((unsigned char)__float2uint_rz(rx), (unsigned char)__float2uint_rz(ry), (unsigned char)__float2uint_rz(rz), 0);
}

// ---------------------------------------------------------------------------------------------------------
// Census + pack.  For every pixel of the PADDED plane: source = clamp(coord) (texture clamp addressing),
// census bit k set iff lum(neighbour_k) > lum(centre), lum = fma(b,.1f, fma(r,.3f, .6f*g)) on RN(k/255) floats
// (bao_pmflow_census_kernel.cu:39-43 as nvcc contracts it), neighbour order TL,T,TR,L,R,BL,B,BR (:53-68).
__device__ __forceinline__ float lum_of(uchar4 p) {
    const float r = __fdiv_rn((float)p.x, 255.f), g = __fdiv_rn((float)p.y, 255.f), b = __fdiv_rn((float)p.z, 255.f);
    return __fmaf_rn(b, 0.1f, __fmaf_rn(r, 0.3f, __fmul_rn(g, 0.6f)));
}

__global__ void __launch_bounds__(256) k_census_pack(const uchar4* __restrict__ rgba, size_t pitch_px, size_t img_stride_px, float4* __restrict__ pix,
                                                     int w, int h, int pw, int ph) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= pw || py >= ph) return;
    const uchar4* src = rgba + (size_t)blockIdx.z * img_stride_px;
    const int x = max(0, min(w - 1, px - PAD)), y = max(0, min(h - 1, py - PAD));
    const int xm = max(0, x - 1), xp = min(w - 1, x + 1), ym = max(0, y - 1), yp = min(h - 1, y + 1);
    const uchar4* r0 = src + (size_t)ym * pitch_px;
    const uchar4* r1 = src + (size_t)y * pitch_px;
    const uchar4* r2 = src + (size_t)yp * pitch_px;
    const uchar4 c = r1[x];
    const float lc = lum_of(c);
    unsigned cen = 0;
    cen |= (lum_of(r0[xm]) > lc) << 0;
    cen |= (lum_of(r0[x]) > lc) << 1;
    cen |= (lum_of(r0[xp]) > lc) << 2;
    cen |= (lum_of(r1[xm]) > lc) << 3;
    cen |= (lum_of(r1[xp]) > lc) << 4;
    cen |= (lum_of(r2[xm]) > lc) << 5;
    cen |= (lum_of(r2[x]) > lc) << 6;
    cen |= (lum_of(r2[xp]) > lc) << 7;
    float4 o;
    o.x = __fdiv_rn((float)c.x, 255.f);
    o.y = __fdiv_rn((float)c.y, 255.f);
    o.z = __fdiv_rn((float)c.z, 255.f);
    o.w = __uint_as_float(pack_census(cen));  // see census_lut
    pix[(size_t)blockIdx.z * pw * ph + (size_t)py * pw + px] = o;
}

// The same with the CTA's 32 x 8 source pixels and their one-pixel halo staged in shared memory by ONE TMA copy (cp.async.bulk.tensor + mbarrier,
// box 40 x 10 uchar4: the 34 x 10 pixels it needs plus three to the left, because the box of a non-interleaved map must start on a 16-byte
// boundary -- an unaligned start is an illegal instruction, tools/probe_tma_u32.cu) and every luminance formed once per tile entry instead of
// nine times per pixel.  A tensor map's out-of-bounds fill is zero,
// not clamp-to-edge, so only CTAs whose tile lies inside the image take this path; the CTAs of the replicated border and of the image rim keep
// the per-thread clamped loads.  Same comparisons on the same floats: the same bits.
__global__ void __launch_bounds__(256) k_census_pack_tile(const uchar4* __restrict__ rgba, size_t pitch_px, size_t img_stride_px, float4* __restrict__ pix,
                                                          int w, int h, int pw, int ph, const __grid_constant__ CUtensorMap tmap) {
    __shared__ __align__(128) uchar4 s_px[CEN_TILE_H][CEN_TILE_W];
    __shared__ float s_lum[CEN_TILE_H][CEN_TILE_W];
    __shared__ __align__(8) unsigned long long s_bar;
    const int px0 = blockIdx.x * 32, py0 = blockIdx.y * 8;
    const int sx0 = px0 - PAD - 1, sy0 = py0 - PAD - 1;   // source pixel of tile entry (0, 0)
    const bool interior = sx0 >= 0 && sy0 >= 0 && sx0 + 34 <= w && sy0 + 10 <= h && px0 + 32 <= pw && py0 + 8 <= ph;   // CTA-uniform
    const int px = px0 + threadIdx.x, py = py0 + threadIdx.y;
    if (interior) {
        const int tid = threadIdx.y * 32 + threadIdx.x;
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            mbar_expect_tx(&s_bar, CEN_TILE_W * CEN_TILE_H * (unsigned)sizeof(uchar4));
            tma_load_3d(&s_px[0][0], &tmap, sx0 - 3, sy0, blockIdx.z, &s_bar);   // sx0 - 3 = px0 - 20: a multiple of 4 pixels
        }
        __syncthreads();   // the barrier object is initialised before anyone polls it
        mbar_wait(&s_bar, 0);
        for (int i = tid; i < (int)(CEN_TILE_H * CEN_TILE_W); i += 256) (&s_lum[0][0])[i] = lum_of((&s_px[0][0])[i]);
        __syncthreads();
        const int tx = threadIdx.x + 4, ty = threadIdx.y + 1;   // tile column of source pixel sx0 + k is k + 3
        const uchar4 c = s_px[ty][tx];
        const float lc = s_lum[ty][tx];
        unsigned cen = 0;
        cen |= (s_lum[ty - 1][tx - 1] > lc) << 0;
        cen |= (s_lum[ty - 1][tx] > lc) << 1;
        cen |= (s_lum[ty - 1][tx + 1] > lc) << 2;
        cen |= (s_lum[ty][tx - 1] > lc) << 3;
        cen |= (s_lum[ty][tx + 1] > lc) << 4;
        cen |= (s_lum[ty + 1][tx - 1] > lc) << 5;
        cen |= (s_lum[ty + 1][tx] > lc) << 6;
        cen |= (s_lum[ty + 1][tx + 1] > lc) << 7;
        float4 o;
        o.x = __fdiv_rn((float)c.x, 255.f);
        o.y = __fdiv_rn((float)c.y, 255.f);
        o.z = __fdiv_rn((float)c.z, 255.f);
        o.w = __uint_as_float(pack_census(cen));
        pix[(size_t)blockIdx.z * pw * ph + (size_t)py * pw + px] = o;
        return;
    }
    if (px >= pw || py >= ph) return;
    const uchar4* src = rgba + (size_t)blockIdx.z * img_stride_px;
    const int x = max(0, min(w - 1, px - PAD)), y = max(0, min(h - 1, py - PAD));
    const int xm = max(0, x - 1), xp = min(w - 1, x + 1), ym = max(0, y - 1), yp = min(h - 1, y + 1);
    const uchar4* r0 = src + (size_t)ym * pitch_px;
    const uchar4* r1 = src + (size_t)y * pitch_px;
    const uchar4* r2 = src + (size_t)yp * pitch_px;
    const uchar4 c = r1[x];
    const float lc = lum_of(c);
    unsigned cen = 0;
    cen |= (lum_of(r0[xm]) > lc) << 0;
    cen |= (lum_of(r0[x]) > lc) << 1;
    cen |= (lum_of(r0[xp]) > lc) << 2;
    cen |= (lum_of(r1[xm]) > lc) << 3;
    cen |= (lum_of(r1[xp]) > lc) << 4;
    cen |= (lum_of(r2[xm]) > lc) << 5;
    cen |= (lum_of(r2[x]) > lc) << 6;
    cen |= (lum_of(r2[xp]) > lc) << 7;
    float4 o;
    o.x = __fdiv_rn((float)c.x, 255.f);
    o.y = __fdiv_rn((float)c.y, 255.f);
    o.z = __fdiv_rn((float)c.z, 255.f);
    o.w = __uint_as_float(pack_census(cen));
    pix[(size_t)blockIdx.z * pw * ph + (size_t)py * pw + px] = o;
}

// census + pack of one of the context's own RGBA levels: the TMA-staged kernel when the level has a tensor map (16-byte row stride)
static void op_census_pack(eppm_context* c, int img, int level, int n_img) {
    const LevelGeom& g = c->lv[level];
    if (c->tmap_rgba_ok[img][level] && !(c->variant & EPPM_VAR_CENSUS_NOTMA)) {
        dim3 blk(32, 8), grd((g.pw + 31) / 32, (g.ph + 7) / 8, n_img);
        k_census_pack_tile<<<grd, blk, 0, c->stream>>>(c->rgba[img][level], (size_t)g.w, (size_t)g.w * g.h, c->pix[img][level], g.w, g.h, g.pw, g.ph,
                                                      c->tmap_rgba[img][level]);
        EPPM_LAUNCH_COUNT(1);
        return;
    }
    k_pack_planes(c->stream, c->rgba[img][level], (size_t)g.w * 4, (size_t)g.w * g.h * 4, c->pix[img][level], g, n_img);
}

void k_pack_planes(cudaStream_t s, const uchar4* rgba, size_t rgba_pitch_bytes, size_t rgba_img_stride_bytes, float4* pix, const LevelGeom& g,
                   int n_img) {
    dim3 blk(32, 8), grd((g.pw + 31) / 32, (g.ph + 7) / 8, n_img);
    k_census_pack<<<grd, blk, 0, s>>>(rgba, rgba_pitch_bytes / 4, rgba_img_stride_bytes / 4, pix, g.w, g.h, g.pw, g.ph);
    EPPM_LAUNCH_COUNT(1);
}

// pack a foreign level (uchar4 rows + u8 census rows, both pitched) into a padded packed plane, census taken as given
__global__ void __launch_bounds__(256) k_pack_foreign(const uchar4* __restrict__ rgba, size_t pitch4, const unsigned char* __restrict__ census,
                                                      size_t pitch1, float4* __restrict__ pix, int w, int h, int pw, int ph) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= pw || py >= ph) return;
    const int x = max(0, min(w - 1, px - PAD)), y = max(0, min(h - 1, py - PAD));
    const uchar4 c = *reinterpret_cast<const uchar4*>(reinterpret_cast<const char*>(rgba) + (size_t)y * pitch4 + (size_t)x * 4);
    const unsigned cen = census ? census[(size_t)y * pitch1 + x] : 0u;
    float4 o;
    o.x = __fdiv_rn((float)c.x, 255.f);
    o.y = __fdiv_rn((float)c.y, 255.f);
    o.z = __fdiv_rn((float)c.z, 255.f);
    o.w = __uint_as_float(pack_census(cen));
    pix[(size_t)py * pw + px] = o;
}

void op_pack_foreign(cudaStream_t s, const uchar4* rgba, size_t rgba_pitch_bytes, const unsigned char* census, size_t census_pitch_bytes, float4* pix,
                     const LevelGeom& g) {
    dim3 blk(32, 8), grd((g.pw + 31) / 32, (g.ph + 7) / 8, 1);
    k_pack_foreign<<<grd, blk, 0, s>>>(rgba, rgba_pitch_bytes, census, census_pitch_bytes, pix, g.w, g.h, g.pw, g.ph);
    EPPM_LAUNCH_COUNT(1);
}

__global__ void k_extract_census_pitched(const float4* __restrict__ pix, int pw, unsigned char* __restrict__ out, size_t pitch, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    out[(size_t)y * pitch + x] = (unsigned char)unpack_census(pix[(size_t)(y + PAD) * pw + x + PAD].w);
}

void op_extract_census(cudaStream_t s, const float4* pix, const LevelGeom& g, unsigned char* out, size_t out_pitch_bytes) {
    k_extract_census_pitched<<<dim3((g.w + 127) / 128, g.h), 128, 0, s>>>(pix, g.pw, out, out_pitch_bytes, g.w, g.h);
    EPPM_LAUNCH_COUNT(1);
}

// Column-major copy of a packed plane (pixel (x,y) at x*ph + y) for the row propagation passes of PatchMatch.
__global__ void __launch_bounds__(256) k_transpose_plane(const float4* __restrict__ src, float4* __restrict__ dst, int pw, int ph) {
    __shared__ float4 tile[16][17];
    const size_t off = (size_t)blockIdx.z * pw * ph;
    int x = blockIdx.x * 16 + threadIdx.x, y = blockIdx.y * 16 + threadIdx.y;
    if (x < pw && y < ph) tile[threadIdx.y][threadIdx.x] = src[off + (size_t)y * pw + x];
    __syncthreads();
    x = blockIdx.x * 16 + threadIdx.y;
    y = blockIdx.y * 16 + threadIdx.x;
    if (x < pw && y < ph) dst[off + (size_t)x * ph + y] = tile[threadIdx.x][threadIdx.y];
}

// Q planes (eppm_device.cuh): every padded pixel goes to its parity sub-plane, in both copies
__global__ void __launch_bounds__(256) k_split_plane(const float4* __restrict__ src, float4* __restrict__ dst, int pw, int ph, QGeom q) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= pw || y >= ph) return;
    const float4 v = src[(size_t)blockIdx.z * pw * ph + (size_t)y * pw + x];
    float4* d = dst + (size_t)blockIdx.z * q.plane;
    const unsigned e = q_index(q, x, y);
    d[e] = v;
    d[e + q.c1] = v;
}

void op_split_plane(cudaStream_t s, const float4* src, float4* dst, const LevelGeom& g, int n_img) {
    dim3 blk(32, 8), grd((g.pw + 31) / 32, (g.ph + 7) / 8, n_img);
    k_split_plane<<<grd, blk, 0, s>>>(src, dst, g.pw, g.ph, make_qgeom(g.pw, g.ph));
    EPPM_LAUNCH_COUNT(1);
}

void op_transpose_plane(cudaStream_t s, const float4* src, float4* dst, const LevelGeom& g, int n_img) {
    dim3 blk(16, 16), grd((g.pw + 15) / 16, (g.ph + 15) / 16, n_img);
    k_transpose_plane<<<grd, blk, 0, s>>>(src, dst, g.pw, g.ph);
    EPPM_LAUNCH_COUNT(1);
}

void op_preblur_rgba(eppm_context* c, const uchar4* src1, const uchar4* src2, size_t pitch_bytes) {
    const LevelGeom& g0 = c->lv[0];
    dim3 blk(PB_T, PB_T), grd((g0.w + PB_T - 1) / PB_T, (g0.h + PB_T - 1) / PB_T, 2);
    k_preblur_rgb<true><<<grd, blk, 0, c->stream>>>(reinterpret_cast<const uint8_t*>(src1), reinterpret_cast<const uint8_t*>(src2), c->rgba[0][0],
                                                  c->rgba[1][0], g0.w, g0.h, c->gauss[0].d_w, pitch_bytes, 2);
    EPPM_LAUNCH_COUNT(1);
}

void run_prepare(eppm_context* c, const uint8_t* d_img1, const uint8_t* d_img2, int n) {
    cudaStream_t s = c->stream;
    const LevelGeom& g0 = c->lv[0];
    {
        dim3 blk(PB_T, PB_T), grd((g0.w + PB_T - 1) / PB_T, (g0.h + PB_T - 1) / PB_T, 2 * n);
        k_preblur_rgb<false><<<grd, blk, 0, s>>>(d_img1, d_img2, c->rgba[0][0], c->rgba[1][0], g0.w, g0.h, c->gauss[0].d_w, 0, 2);
        EPPM_LAUNCH_COUNT(1);
    }
    op_pyramid_and_pack(c, n);
}

// Video stream: n_frames consecutive frames, each prepared ONCE into the image-1 arrays (frame f = plane f); pair f then reads
// frame f as image 1 and frame f+1 as image 2 through base pointers one plane apart (run_stream in context.cu).
void run_prepare_frames(eppm_context* c, const uint8_t* d_frames, int n_frames) {
    const LevelGeom& g0 = c->lv[0];
    dim3 blk(PB_T, PB_T), grd((g0.w + PB_T - 1) / PB_T, (g0.h + PB_T - 1) / PB_T, n_frames);
    k_preblur_rgb<false><<<grd, blk, 0, c->stream>>>(d_frames, d_frames, c->rgba[0][0], c->rgba[0][0], g0.w, g0.h, c->gauss[0].d_w, 0, 1);
    EPPM_LAUNCH_COUNT(1);
    op_pyramid_and_pack(c, n_frames, 1);
}

void op_pyramid_and_pack(eppm_context* c, int n, int two) {
    cudaStream_t s = c->stream;
    const LevelGeom& g0 = c->lv[0];
    for (int i = 1; i < c->n_levels; i++) {
        const LevelGeom& g = c->lv[i];
        const LevelGeom& gs = c->lv[i - 1];  // source level: i-n with n = 1
        const int r = c->gauss[i].r;
        // :656 ratio = pow(ratio,i) for i <= n (= 0.5 at i = 1); :661 (float)pow(ratio,i)*W[0]/W[i-n] beyond
        const float ratio = i == 1 ? 0.5f : (float)pow((double)0.5f, i) * g0.w / gs.w;
        if (ratio == 0.5f) {
            dim3 blk(128), grd((g.w + 127) / 128, g.h, two * n);
            size_t smem = ((2 * r + 1) * (2 * r + 1) + 1) * sizeof(float);
            k_pyr_decimate<<<grd, blk, smem, s>>>(c->rgba[0][i - 1], c->rgba[1][i - 1], c->rgba[0][i], c->rgba[1][i], gs.w, gs.h, g.w, g.h, 2, r,
                                                  c->gauss[i].d_w, two);
            EPPM_LAUNCH_COUNT(1);
        } else {
            for (int img = 0; img < two; img++) {
                dim3 blk(128), grd((gs.w + 127) / 128, gs.h, n);
                k_blur_full<<<grd, blk, 0, s>>>(c->rgba[img][i - 1], c->blur_tmp[img], gs.w, gs.h, r, c->gauss[i].d_w);
                dim3 grd2((g.w + 127) / 128, g.h, n);
                k_resize_u8<<<grd2, blk, 0, s>>>(c->rgba[img][i], g.w, g.h, c->blur_tmp[img], gs.w, gs.h, ratio);
                EPPM_LAUNCH_COUNT(2);
            }
        }
    }
    for (int i = 0; i < c->n_levels; i++) {
        const LevelGeom& g = c->lv[i];
        for (int img = 0; img < two; img++) op_census_pack(c, img, i, n);
    }
    const int L = c->n_levels - 1;
    for (int img = 0; img < two; img++) op_transpose_plane(s, c->pix[img][L], c->pixT[img], c->lv[L], n);
    if (c->pixQ[0])
        for (int img = 0; img < two; img++) op_split_plane(s, c->pix[img][L], c->pixQ[img], c->lv[L], n);
}

}  // namespace eppm
