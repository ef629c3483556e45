// Optional final stage of the reference's C ABI (not called by compute_flow): true sub-pixel refinement.
//   baoCudaCensusTransform_Bicubic  (bao_pmflow_census_kernel.cu:115-181)  census of the x2 bicubic (B-spline) upsampled images
//   baoCudaSubpixRefine             (bao_pmflow_refine_kernel.cu:395-634, 678-722)
// For every pixel with a valid integer target D: the bilateral AD+census patch cost is sampled on the 5x5 half-pixel
// neighbourhood of D (100 half-pixel spaced samples per cost, colours by bicubic filtering, census from the upsampled planes), a
// quadric a x^2 + b y^2 + c xy + d x + e y + f is fitted by least squares (normal equations, <= 5 conjugate-gradient steps) and its
// stationary point, if within 3 half-pixels, replaces the integer flow.
//
// B200 design (the reference runs one thread per pixel and re-filters every colour it touches: 25 x 100 x 2 bicubic look-ups of
// 16 texture fetches each = 80 000 fetches per pixel):
//  * one WARP per pixel.  All colours a pixel needs lie on a half-pixel grid: 10x10 sites around the pixel in image 1 and 23x23
//    around D in image 2.  The warp fetches the 13x13 / 15x15 source samples once (texture unit, see below), filters rows then
//    columns into shared memory (each row result feeds up to four column filters) and keeps the tables for all 25 costs:
//    394 fetches and 1 200 cubic filters per pixel instead of 80 000 and 12 500.
//  * lane p < 25 owns neighbourhood position p and adds its 100 samples in the reference's order; the 6x6 solve runs once per warp.
//  * the image textures are bound exactly like the reference binds them: pitch-2D, normalised-float reads, LINEAR filtering, and
//    fetched at integer coordinates -- i.e. half way between texels, so every "texel" the bicubic filter sees is the texture unit's
//    own 2x2 average.  Using the same unit on the same coordinates gives the same bits; the arithmetic after it is spelled out
//    with round-to-nearest intrinsics in the contraction nvcc 12.9 produces for the reference (read from its PTX and SASS).
#include <float.h>
#include <math.h>
#include <stdio.h>

#include "../../include/eppm_legacy_abi.h"
#include "eppm_internal.h"

namespace eppm {

// B-spline weights w0..w3 of bicubicTexture_kernel.cuh:33-57 as contracted:
//   w0 = (1/6) fma(a, fma(a, 3 - a, -3), 1)        w1 = (1/6) fma(a*a, fma(a, 3, -6), 4)
//   w2 = (1/6) fma(a, fma(a, fma(a, -3, 3), 3), 1)  w3 = (1/6) (a * (a*a))
__device__ __forceinline__ void bspline_weights(float a, float (&w)[4]) {
    const float sixth = 0.16666667163372039795f;
    const float a2 = __fmul_rn(a, a);
    w[0] = __fmul_rn(__fmaf_rn(a, __fmaf_rn(a, __fsub_rn(3.0f, a), -3.0f), 1.0f), sixth);
    w[1] = __fmul_rn(__fmaf_rn(a2, __fmaf_rn(a, 3.0f, -6.0f), 4.0f), sixth);
    w[2] = __fmul_rn(__fmaf_rn(a, __fmaf_rn(a, __fmaf_rn(a, -3.0f, 3.0f), 3.0f), 1.0f), sixth);
    w[3] = __fmul_rn(__fmul_rn(a, a2), sixth);
}
// cubicFilter (:75-84): r = c0*w0; r += c1*w1; r += c2*w2; r += c3*w3  ->  fma(w3,c3, fma(w2,c2, fma(w0,c0, w1*c1)))
__device__ __forceinline__ float cubic4(const float (&w)[4], float c0, float c1, float c2, float c3) {
    return __fmaf_rn(w[3], c3, __fmaf_rn(w[2], c2, __fmaf_rn(w[0], c0, __fmul_rn(w[1], c1))));
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Census of the bicubic-upsampled image (point-sampled texture in the reference: plain texel reads with clamp addressing).
// V(X, Y) = bicubic(img, X*up, Y*up); bit k of census(X, Y) = lum(V(neighbour k)) > lum(V(X, Y)).  A CTA computes lum(V) once for its
// 32x8 tile plus a one-pixel halo (the reference evaluates nine bicubic look-ups per pixel and image).
__device__ __forceinline__ float unorm8(unsigned char v) { return __fdiv_rn((float)v, 255.f); }

__device__ float bicubic_lum_point(const uchar4* __restrict__ img, size_t pitch_px, int w, int h, int X, int Y, float up) {
    // tex2DBicubic (:88-104): x -= 0.5; px = floor(x); fx = x - px; 16 point fetches at (px-1..px+2, py-1..py+2)
    const float x = __fsub_rn(__fmul_rn((float)X, up), 0.5f), y = __fsub_rn(__fmul_rn((float)Y, up), 0.5f);
    const float px = floorf(x), py = floorf(y);
    float wx[4], wy[4];
    bspline_weights(__fsub_rn(x, px), wx);
    bspline_weights(__fsub_rn(y, py), wy);
    const int ix = (int)px, iy = (int)py;
    float row[4][3];
#pragma unroll
    for (int l = 0; l < 4; l++) {
        const int cy = max(0, min(h - 1, iy - 1 + l));
        uchar4 t[4];
#pragma unroll
        for (int k = 0; k < 4; k++) t[k] = img[(size_t)cy * pitch_px + max(0, min(w - 1, ix - 1 + k))];
        row[l][0] = cubic4(wx, unorm8(t[0].x), unorm8(t[1].x), unorm8(t[2].x), unorm8(t[3].x));
        row[l][1] = cubic4(wx, unorm8(t[0].y), unorm8(t[1].y), unorm8(t[2].y), unorm8(t[3].y));
        row[l][2] = cubic4(wx, unorm8(t[0].z), unorm8(t[1].z), unorm8(t[2].z), unorm8(t[3].z));
    }
    const float r = cubic4(wy, row[0][0], row[1][0], row[2][0], row[3][0]);
    const float g = cubic4(wy, row[0][1], row[1][1], row[2][1], row[3][1]);
    const float b = cubic4(wy, row[0][2], row[1][2], row[2][2], row[3][2]);
    return __fmaf_rn(b, 0.1f, __fmaf_rn(r, 0.3f, __fmul_rn(g, 0.6f)));   // _d_is_larger (:39-43) as contracted
}

constexpr int CB_TX = 32, CB_TY = 8;
__global__ void __launch_bounds__(CB_TX* CB_TY) k_census_bicubic(unsigned char* __restrict__ cen1, unsigned char* __restrict__ cen2, int w_up, int h_up,
                                                                 size_t census_w, const uchar4* __restrict__ img1, const uchar4* __restrict__ img2,
                                                                 size_t pitch_px, int w, int h, float up) {
    __shared__ float s_lum[CB_TY + 2][CB_TX + 2];
    const int X0 = blockIdx.x * CB_TX, Y0 = blockIdx.y * CB_TY;
    const int tid = threadIdx.y * CB_TX + threadIdx.x;
    const uchar4* img = blockIdx.z ? img2 : img1;
    unsigned char* cen = blockIdx.z ? cen2 : cen1;
    for (int i = tid; i < (CB_TY + 2) * (CB_TX + 2); i += CB_TX * CB_TY) {
        const int ty = i / (CB_TX + 2), tx = i - ty * (CB_TX + 2);
        s_lum[ty][tx] = bicubic_lum_point(img, pitch_px, w, h, X0 + tx - 1, Y0 + ty - 1, up);   // ids -1 and w_up are evaluated, not clamped
    }
    __syncthreads();
    const int X = X0 + threadIdx.x, Y = Y0 + threadIdx.y;
    if (X >= w_up || Y >= h_up) return;
    const int tx = threadIdx.x + 1, ty = threadIdx.y + 1;
    const float c = s_lum[ty][tx];
    unsigned r = 0;   // neighbour order TL,T,TR,L,R,BL,B,BR (:123-139)
    r |= (s_lum[ty - 1][tx - 1] > c) << 0;
    r |= (s_lum[ty - 1][tx] > c) << 1;
    r |= (s_lum[ty - 1][tx + 1] > c) << 2;
    r |= (s_lum[ty][tx - 1] > c) << 3;
    r |= (s_lum[ty][tx + 1] > c) << 4;
    r |= (s_lum[ty + 1][tx - 1] > c) << 5;
    r |= (s_lum[ty + 1][tx] > c) << 6;
    r |= (s_lum[ty + 1][tx + 1] > c) << 7;
    cen[(size_t)Y * census_w + X] = (unsigned char)r;
}

// ---------------------------------------------------------------------------------------------------------------------------------
struct SubpixLut {
    float g[10];        // expf(-i^2 / SUBPIX_SIG_S^2), SUBPIX_SIG_S = 9 (:701-705)
    float census[9];    // 1 - expf(-i^2 / (LAMBDA_CENSUS*8)^2) (:707-711)
    float ata[6][6];    // A^T A of the 25 x 6 design matrix (x^2, y^2, xy, x, y, 1), x, y in -2..2 (:580-606): small integers
};

struct SubpixArgs {
    float2* flow; size_t flow_w;
    const short2* nnf; size_t disp_w;
    cudaTextureObject_t img1, img2;              // uchar4, normalised float, LINEAR, clamp, unnormalised coordinates
    const unsigned char* cen1; const unsigned char* cen2; size_t census_w;   // census of the x2 upsampled images, [2h][2w]
    int w, h;
};

constexpr int SP_WARPS = 4;
constexpr int SP_F = 15, SP_G = 23;   // source samples per axis around D; half-pixel sites per axis around D
// per-warp shared memory (floats): F 3x225, R 3x(15x23), B2 3x529, B1 3x100, vecB 32, then census bytes C2 529 + C1 100
constexpr int SP_OFF_F = 0, SP_OFF_R = SP_OFF_F + 3 * SP_F * SP_F, SP_OFF_B2 = SP_OFF_R + 3 * SP_F * SP_G, SP_OFF_B1 = SP_OFF_B2 + 3 * SP_G * SP_G,
              SP_OFF_VB = SP_OFF_B1 + 300, SP_OFF_C = SP_OFF_VB + 32, SP_WARP_FLOATS = SP_OFF_C + (SP_G * SP_G + 100 + 3) / 4 + 1;
constexpr size_t SP_SMEM = (size_t)SP_WARPS * SP_WARP_FLOATS * sizeof(float);

__device__ __forceinline__ float3 tex_rgb(cudaTextureObject_t t, float x, float y) {
    const float4 v = tex2D<float4>(t, x, y);
    return make_float3(v.x, v.y, v.z);
}
__device__ __forceinline__ float max3abs(float ax, float ay, float az, float bx, float by, float bz) {
    return fmaxf(fmaxf(fabsf(__fsub_rn(ax, bx)), fabsf(__fsub_rn(ay, by))), fabsf(__fsub_rn(az, bz)));
}

__global__ void __launch_bounds__(SP_WARPS * 32) k_subpix_refine(SubpixArgs a, const __grid_constant__ SubpixLut lut) {
    extern __shared__ float sp_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* S = sp_smem + warp * SP_WARP_FLOATS;
    float* F = S + SP_OFF_F;      // [c][b][a]
    float* R = S + SP_OFF_R;      // [c][b][u]
    float* B2 = S + SP_OFF_B2;    // [c][v][u]
    float* B1 = S + SP_OFF_B1;    // [c][ii][jj]
    float* VB = S + SP_OFF_VB;
    unsigned char* C2 = reinterpret_cast<unsigned char*>(S + SP_OFF_C);
    unsigned char* C1 = C2 + SP_G * SP_G;
    const int x1 = blockIdx.x * SP_WARPS + warp, y1 = blockIdx.y;
    if (x1 >= a.w) return;                                                   // whole warp
    const short2 D = a.nnf[(size_t)y1 * a.disp_w + x1];
    if (D.x < 0 || D.y < 0 || D.x >= a.w || D.y >= a.h) return;              // :545 (flow untouched)
    const float fx1 = (float)x1, fy1 = (float)y1, fDx = (float)D.x, fDy = (float)D.y;
    float w0[4], wh[4];
    bspline_weights(0.0f, w0);   // the fractional part is 0 or 0.5 on the half-pixel grid
    bspline_weights(0.5f, wh);

    // ---- image 1: sites (x1 + j/2, y1 + i/2), i, j odd in -9..9: px = x1 - 5 + jj, fraction 0; samples x1-6 .. x1+6
    for (int k = lane; k < 13 * 13; k += 32) {
        const int b = k / 13, aa = k - b * 13;
        const float3 v = tex_rgb(a.img1, fx1 + (float)(aa - 6), fy1 + (float)(b - 6));
        F[k] = v.x; F[225 + k] = v.y; F[450 + k] = v.z;
    }
    for (int k = lane; k < 100; k += 32) {                                   // census1 at ((x1 + j/2)*2, (y1 + i/2)*2) = (2 x1 + j, 2 y1 + i), clamped (:464)
        const int ii = k / 10, jj = k - ii * 10;
        const int cx = max(0, min(2 * a.w - 1, 2 * x1 - 9 + 2 * jj)), cy = max(0, min(2 * a.h - 1, 2 * y1 - 9 + 2 * ii));
        C1[k] = a.cen1[(size_t)cy * a.census_w + cx];
    }
    __syncwarp();
    for (int k = lane; k < 3 * 13 * 10; k += 32) {                           // row filters
        const int c = k / 130, r = k - c * 130, b = r / 10, jj = r - b * 10;
        const float* f = F + c * 225 + b * 13 + jj;
        R[k] = cubic4(w0, f[0], f[1], f[2], f[3]);
    }
    __syncwarp();
    for (int k = lane; k < 300; k += 32) {                                   // column filters
        const int c = k / 100, r = k - c * 100, ii = r / 10, jj = r - ii * 10;
        const float* q = R + c * 130 + ii * 10 + jj;
        B1[k] = cubic4(w0, q[0], q[10], q[20], q[30]);
    }
    const float3 c1 = tex_rgb(a.img1, fx1, fy1);                             // centerPix1 (:443)
    __syncwarp();

    // ---- image 2: sites (D.x + u/2, D.y + v/2), u, v in -11..11: px = D.x + floor((u-1)/2), fraction ((u-1) & 1)/2; samples D-7 .. D+7
    for (int k = lane; k < SP_F * SP_F; k += 32) {
        const int b = k / SP_F, aa = k - b * SP_F;
        const float3 v = tex_rgb(a.img2, fDx + (float)(aa - 7), fDy + (float)(b - 7));
        F[k] = v.x; F[225 + k] = v.y; F[450 + k] = v.z;
    }
    for (int k = lane; k < SP_G * SP_G; k += 32) {                           // census2 at (2 D.x + u, 2 D.y + v), clamped (:465)
        const int vv = k / SP_G, uu = k - vv * SP_G;
        const int cx = max(0, min(2 * a.w - 1, 2 * D.x + uu - 11)), cy = max(0, min(2 * a.h - 1, 2 * D.y + vv - 11));
        C2[k] = a.cen2[(size_t)cy * a.census_w + cx];
    }
    __syncwarp();
    for (int k = lane; k < 3 * SP_F * SP_G; k += 32) {
        const int c = k / (SP_F * SP_G), r = k - c * (SP_F * SP_G), b = r / SP_G, uu = r - b * SP_G;
        const int um1 = uu - 12;                                             // u - 1
        const int a0 = (um1 >> 1) + 6;                                       // px - 1 - (D.x - 7), floor division by the arithmetic shift
        const float* f = F + c * 225 + b * SP_F + a0;
        R[k] = (um1 & 1) ? cubic4(wh, f[0], f[1], f[2], f[3]) : cubic4(w0, f[0], f[1], f[2], f[3]);
    }
    __syncwarp();
    for (int k = lane; k < 3 * SP_G * SP_G; k += 32) {
        const int c = k / (SP_G * SP_G), r = k - c * (SP_G * SP_G), vv = r / SP_G, uu = r - vv * SP_G;
        const int vm1 = vv - 12;
        const int b0 = (vm1 >> 1) + 6;
        const float* q = R + c * (SP_F * SP_G) + b0 * SP_G + uu;
        B2[k] = (vm1 & 1) ? cubic4(wh, q[0], q[SP_G], q[2 * SP_G], q[3 * SP_G]) : cubic4(w0, q[0], q[SP_G], q[2 * SP_G], q[3 * SP_G]);
    }
    __syncwarp();

    // ---- the 25 costs: lane p = (dy + 2) * 5 + dx + 2  (:548-567)
    float vb = 0.f;
    bool valid = false;
    if (lane < 25) {
        const int dy = lane / 5 - 2, dx = lane - (lane / 5) * 5 - 2;
        const float nx = __fmaf_rn((float)dx, 0.5f, fDx), ny = __fmaf_rn((float)dy, 0.5f, fDy);
        valid = !(nx < 0.f || nx >= (float)a.w || ny < 0.f || ny >= (float)a.h);
        vb = 2.0f;
        if (valid) {
            const float3 c2 = tex_rgb(a.img2, nx, ny);                       // centerPix2 (:444)
            float cs = 0.f, ws = 0.f;
            for (int ii = 0; ii < 10; ii++) {
                const float gi = lut.g[abs(2 * ii - 9)];
                const int rowb = (dy + 2 * ii - 9 + 11) * SP_G + dx + 11 - 9;
                for (int jj = 0; jj < 10; jj++) {
                    const int s = ii * 10 + jj, t = rowb + 2 * jj;
                    const float ar = B1[s], ag = B1[100 + s], ab = B1[200 + s];
                    const float br = B2[t], bg = B2[SP_G * SP_G + t], bb = B2[2 * SP_G * SP_G + t];
                    // _d_subpix_bilateral_dist (:402-421)
                    const float mod = max3abs(br, bg, bb, ar, ag, ab);
                    const float ad = __fadd_rn(1.0f, -exp_ref(__fdiv_rn(__fmul_rn(mod, mod), -0.010000000707805156708f)));
                    const float cen = lut.census[__popc((unsigned)(C1[s] ^ C2[t]))];
                    const float d1 = max3abs(c1.x, c1.y, c1.z, ar, ag, ab), d2 = max3abs(c2.x, c2.y, c2.z, br, bg, bb);
                    const float coef_r = exp_ref(__fdiv_rn(__fmaf_rn(d1, d1, __fmul_rn(d2, d2)), -0.040000002831220626831f));
                    const float wgt = __fmul_rn(coef_r, __fmul_rn(lut.g[abs(2 * jj - 9)], gi));
                    cs = __fmaf_rn(wgt, __fadd_rn(ad, cen), cs);
                    ws = __fadd_rn(wgt, ws);
                }
            }
            vb = __fdiv_rn(cs, ws);
        }
        VB[lane] = vb;
    }
    if (!__any_sync(0xffffffffu, valid)) return;                             // :568
    __syncwarp();

    // ---- least squares: A^T b in equation order, conjugate gradient on A^T A (:570-617, 473-536); every lane computes the same
    float atb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int e = 0; e < 25; e++) {
        const float fx = (float)(e % 5 - 2), fy = (float)(e / 5 - 2), b = VB[e];
        atb[0] = __fmaf_rn(fx * fx, b, atb[0]);
        atb[1] = __fmaf_rn(fy * fy, b, atb[1]);
        atb[2] = __fmaf_rn(fx * fy, b, atb[2]);
        atb[3] = __fmaf_rn(fx, b, atb[3]);
        atb[4] = __fmaf_rn(fy, b, atb[4]);
        atb[5] = __fmaf_rn(1.0f, b, atb[5]);
    }
    float X[6], r[6], d[6], ad[6];
    float nb = 0.f;
#pragma unroll
    for (int i = 0; i < 6; i++) nb = __fmaf_rn(atb[i], atb[i], nb);
    const float normb = __fsqrt_rn(nb);
#pragma unroll
    for (int i = 0; i < 6; i++) { X[i] = 0.f; r[i] = atb[i]; d[i] = atb[i]; }
    float rtr = __fmul_rn(normb, normb);
    int it = 0;
    while ((double)__fdiv_rn(__fsqrt_rn(rtr), normb) > 1.0e-6 && it < 5) {   // the tolerance is a double literal (:496)
        it++;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 6; j++) acc = __fmaf_rn(lut.ata[i][j], d[j], acc);
            ad[i] = acc;
        }
        float dad = 0.f;
#pragma unroll
        for (int i = 0; i < 6; i++) dad = __fmaf_rn(d[i], ad[i], dad);
        const float alpha = __fdiv_rn(rtr, dad);
#pragma unroll
        for (int i = 0; i < 6; i++) {
            X[i] = __fmaf_rn(alpha, d[i], X[i]);
            r[i] = __fmaf_rn(ad[i], -alpha, r[i]);
        }
        const float rtrold = rtr;
        rtr = 0.f;
#pragma unroll
        for (int i = 0; i < 6; i++) rtr = __fmaf_rn(r[i], r[i], rtr);
        const float beta = __fdiv_rn(rtr, rtrold);
#pragma unroll
        for (int i = 0; i < 6; i++) d[i] = __fmaf_rn(beta, d[i], r[i]);
    }
    // stationary point of the quadric (:619-633)
    const float den = __fmaf_rn(X[2], X[2], __fmul_rn(__fmul_rn(X[0], -4.0f), X[1]));
    if (den == 0.f) return;
    const float subx = __fdiv_rn(__fmaf_rn(__fadd_rn(X[3], X[3]), X[1], -__fmul_rn(X[4], X[2])), den);
    const float suby = __fdiv_rn(__fmaf_rn(__fadd_rn(X[0], X[0]), X[4], -__fmul_rn(X[3], X[2])), den);
    if (fabsf(suby) <= 3.0f && fabsf(subx) <= 3.0f && lane == 0) {   // NaN fails both, like the reference's compare
        float2 o;
        o.x = __fmul_rn(__fmaf_rn((float)(D.x - x1), 2.0f, subx), 0.5f);
        o.y = __fmul_rn(__fmaf_rn((float)(D.y - y1), 2.0f, suby), 0.5f);
        a.flow[(size_t)y1 * a.flow_w + x1] = o;
    }
}

static bool make_linear_tex(cudaTextureObject_t* t, const uchar4* img, int w, int h, size_t pitch) {
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = const_cast<uchar4*>(img);
    rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
    rd.res.pitch2D.width = w;
    rd.res.pitch2D.height = h;
    rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;            // :682-683
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 0;
    return cuda_ok(cudaCreateTextureObject(t, &rd, &td, nullptr), "cudaCreateTextureObject(subpix)");
}

}  // namespace eppm

using namespace eppm;

extern "C" {

void baoCudaCensusTransform_Bicubic(unsigned char* d_census1, unsigned char* d_census2, int w_up, int h_up, size_t census_pitch, uchar4* d_img1,
                                    uchar4* d_img2, int w, int h, size_t img_pitch) {
    cudaStreamSynchronize(0);
    const float up = (float)w / (float)w_up;   // :173
    dim3 blk(CB_TX, CB_TY), grd((w_up + CB_TX - 1) / CB_TX, (h_up + CB_TY - 1) / CB_TY, 2);
    k_census_bicubic<<<grd, blk>>>(d_census1, d_census2, w_up, h_up, census_pitch, d_img1, d_img2, img_pitch / sizeof(uchar4), w, h, up);
    EPPM_LAUNCH_COUNT(1);
    if (!cuda_ok(cudaStreamSynchronize(0), "baoCudaCensusTransform_Bicubic")) fprintf(stderr, "EPPM(b200) baoCudaCensusTransform_Bicubic: %s\n", eppm_last_error());
}

void baoCudaSubpixRefine(float2* d_flow, short2* d_disp_vec, uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1_up, unsigned char* d_census2_up,
                         int w, int h, size_t img_pitch, size_t census_pitch_up, size_t disp_pitch, size_t flow_pitch) {
    cudaStreamSynchronize(0);
    SubpixArgs a;
    a.flow = d_flow; a.flow_w = flow_pitch / sizeof(float2);
    a.nnf = d_disp_vec; a.disp_w = disp_pitch / sizeof(short2);
    a.cen1 = d_census1_up; a.cen2 = d_census2_up; a.census_w = census_pitch_up;
    a.w = w; a.h = h;
    a.img1 = a.img2 = 0;
    if (!make_linear_tex(&a.img1, d_img1, w, h, img_pitch) || !make_linear_tex(&a.img2, d_img2, w, h, img_pitch)) {
        fprintf(stderr, "EPPM(b200) baoCudaSubpixRefine: %s\n", eppm_last_error());
        if (a.img1) cudaDestroyTextureObject(a.img1);
        return;
    }
    SubpixLut lut;
    volatile float sig = 9.0f, lc = 0.3f;      // SUBPIX_SIG_S, LAMBDA_CENSUS (defs.h:75,52); volatile: evaluated at run time like the reference
    for (int i = 0; i < 10; i++) lut.g[i] = expf(-float(i * i) / (sig * sig));
    for (int i = 0; i < 9; i++) lut.census[i] = 1 - expf(-float(i * i) / (lc * 8 * lc * 8));
    for (int p = 0; p < 6; p++)
        for (int q = 0; q < 6; q++) {
            int s = 0;
            for (int e = 0; e < 25; e++) {
                const int x = e % 5 - 2, y = e / 5 - 2;
                const int col[6] = {x * x, y * y, x * y, x, y, 1};
                s += col[p] * col[q];
            }
            lut.ata[p][q] = (float)s;
        }
    cudaFuncSetAttribute(k_subpix_refine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SP_SMEM);   // per device; cheap next to the stage
    dim3 grd((w + SP_WARPS - 1) / SP_WARPS, h);
    k_subpix_refine<<<grd, SP_WARPS * 32, SP_SMEM>>>(a, lut);
    EPPM_LAUNCH_COUNT(1);
    if (!cuda_ok(cudaStreamSynchronize(0), "baoCudaSubpixRefine")) fprintf(stderr, "EPPM(b200) baoCudaSubpixRefine: %s\n", eppm_last_error());
    cudaDestroyTextureObject(a.img1);
    cudaDestroyTextureObject(a.img2);
}

}  // extern "C"

// eppm_params::subpixel_final (opt-in): the two stage functions above applied to the context's own level-0 planes, pair by pair, between the level-0
// refine (integer flow in flow_tmp) and the level-0 smoothing passes.  The stage functions work on the legacy default stream and synchronise,
// like the reference's; this path is an accuracy option, not part of the timed pipeline.
extern "C" void baoCudaFlow2NNF(short2* d_disp_vec, float2* d_flow, int w, int h, size_t disp_pitch, size_t flow_pitch);
namespace eppm {
void op_subpix_final(eppm_context* c) {
    const LevelGeom& g = c->lv[0];
    const size_t n = (size_t)g.w * g.h;
    if (!c->subpix_census[0]) {
        for (int i = 0; i < 2; i++)
            if (!cuda_ok(cudaMalloc((void**)&c->subpix_census[i], 4 * n), "cudaMalloc(subpix census)")) return;
        if (!cuda_ok(cudaMalloc((void**)&c->subpix_nnf, n * sizeof(short2)), "cudaMalloc(subpix nnf)")) return;
    }
    cudaStreamSynchronize(c->stream);
    for (int b = 0; b < c->n_cur; b++) {
        uchar4* i1 = c->rgba[0][0] + (size_t)b * n;
        uchar4* i2 = c->rgba[1][0] + (size_t)b * n;
        float2* fl = c->flow_tmp + (size_t)b * n;
        baoCudaCensusTransform_Bicubic(c->subpix_census[0], c->subpix_census[1], 2 * g.w, 2 * g.h, (size_t)2 * g.w, i1, i2, g.w, g.h, (size_t)g.w * 4);
        baoCudaFlow2NNF(c->subpix_nnf, fl, g.w, g.h, (size_t)g.w * sizeof(short2), (size_t)g.w * sizeof(float2));
        baoCudaSubpixRefine(fl, c->subpix_nnf, i1, i2, c->subpix_census[0], c->subpix_census[1], g.w, g.h, (size_t)g.w * 4, (size_t)2 * g.w,
                            (size_t)g.w * sizeof(short2), (size_t)g.w * sizeof(float2));
    }
}
}  // namespace eppm
