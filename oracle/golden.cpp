// TEST INFRASTRUCTURE ONLY — see golden.h for scope, parity status and who may call this.
//
// Single-threaded CPU restatement of EPPM's dense-correspondence path.  Every function cites the reference
// file:line (relative to linchaobao/EPPM) whose behaviour it restates.  Written for clarity, not speed: plain loops over
// pixels, float arithmetic spelled with fmaf() where nvcc contracts the reference's expressions (the contraction pattern
// was read from the SASS of the reference build), -ffp-contract=off so the host compiler adds none.
#include "golden.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "golden_pinned.h"

namespace {

struct U4 { uint8_t x, y, z, w; };
struct S2 { int16_t x, y; };
struct F2 { float x, y; };
struct F3 { float x, y, z; };

const int PATCH_R = 9;                 // defs.h:42
const int INVALID_LOCATION = -10000;   // bao_pmflow_refine_kernel.cu:46
const float UNKNOWN_FLOW = 1e10f, UNKNOWN_FLOW_THRESH = 1e9f;  // defs.h:84-91

inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// __expf(x) = ex2.approx(x * log2e) on the device; the host can only approximate MUFU.EX2 (parity: tolerance-based)
inline float expf_dev(float x) { return exp2f(x * 1.4426950216293334961f); }

struct Level {
    int w, h;
    std::vector<U4> rgba[2];
    std::vector<uint8_t> census[2];
    std::vector<F3> col[2];   // RN(k/255) floats, what a cudaReadModeNormalizedFloat fetch returns (probed on B200)
    std::vector<F2> flow;
    // clamp-addressed reads (texture clamp mode of the reference)
    const F3& C(int img, int x, int y) const { return col[img][(size_t)clampi(y, 0, h - 1) * w + clampi(x, 0, w - 1)]; }
    uint8_t Cen(int img, int x, int y) const { return census[img][(size_t)clampi(y, 0, h - 1) * w + clampi(x, 0, w - 1)]; }
};

inline float max3abs(const F3& a, const F3& b) {
    float dx = fabsf(a.x - b.x), dy = fabsf(a.y - b.y), dz = fabsf(a.z - b.z);
    return fmaxf(fmaxf(dx, dy), dz);
}

}  // namespace

struct golden_ctx {
    int n_levels, num_iter;
    int num_guess = 6, search_range = 30, radius_min = 1, seg_len = 10;  // defs.h:36-38, bao_pmflow_kernel.cu:979
    int stride = 2;  // sample stride of the patch loops (bao_pmflow_kernel.cu:269,272)
    int pf_cost = 0; // 1: the forward PatchMatch scores with the plane-fitting cost (baoCudaPatchMatch_PlaneFitting, bao_pmflow_kernel.cu:1897-1963)
    std::vector<Level> lv;
    std::vector<S2> nnf[2], rng_init, rng_search;
    std::vector<float> cost[2];
    std::vector<float> scale, rng_init_scale, rng_search_scale;   // baoCudaPatchMatch_Scaled only (patchmatch_scaled)
    float G[PATCH_R + 1], census_lut[9], wmf_g[5], blf_g[11];
};

namespace {

// ------------------------------------------------------------------------------------------------ geometry
// bao_pyr_init_dim (basic/bao_basic.h:196-211): arr[i] = int(double(dim) * pow(double(0.5f), i))
void level_dims(int h, int w, int level, int* oh, int* ow) {
    *oh = level == 0 ? h : (int)((double)h * pow((double)0.5f, level));
    *ow = level == 0 ? w : (int)((double)w * pow((double)0.5f, level));
}

// ------------------------------------------------------------------------------------------------ prepare
// _d_bao_gauss_filter<uchar4> (basic/bao_basic_cuda.cuh:437-467) at one output site: clamped (2r+1)^2 taps, dy outer / dx
// inner, val = fma(weight, float(u8), val), sum += weight, result = trunc(val / sum).  Weights: pinned device values.
U4 blur_at(const std::vector<U4>& src, int w, int h, int x, int y, int r, const uint32_t* wbits) {
    float vx = 0.f, vy = 0.f, vz = 0.f, sum = 0.f;
    const int n = 2 * r + 1;
    for (int dy = -r; dy <= r; dy++)
        for (int dx = -r; dx <= r; dx++) {
            const U4& p = src[(size_t)clampi(y + dy, 0, h - 1) * w + clampi(x + dx, 0, w - 1)];
            const float wg = bits2f(wbits[(dy + r) * n + dx + r]);
            vx = fmaf(wg, (float)p.x, vx);
            vy = fmaf(wg, (float)p.y, vy);
            vz = fmaf(wg, (float)p.z, vz);
            sum = sum + wg;
        }
    U4 o;
    o.x = (uint8_t)(unsigned)(vx / sum);
    o.y = (uint8_t)(unsigned)(vy / sum);
    o.z = (uint8_t)(unsigned)(vz / sum);
    o.w = 0;
    return o;
}

// d_census_transform3x3 (bao_pmflow_census_kernel.cu:39-90): lum = fma(b,.1f, fma(r,.3f, .6f*g)) as nvcc contracts
// 0.3f*r + 0.6f*g + 0.1f*b; bit k set iff lum(neighbour k) > lum(centre); order TL,T,TR,L,R,BL,B,BR; clamp addressing.
inline float lum(const F3& c) { return fmaf(c.z, 0.1f, fmaf(c.x, 0.3f, c.y * 0.6f)); }

void census_level(Level& L, int img) {
    static const int ox[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, oy[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
    L.census[img].resize((size_t)L.w * L.h);
    for (int y = 0; y < L.h; y++)
        for (int x = 0; x < L.w; x++) {
            const float lc = lum(L.C(img, x, y));
            unsigned c = 0;
            for (int k = 0; k < 8; k++) c |= (unsigned)(lum(L.C(img, x + ox[k], y + oy[k])) > lc) << k;
            L.census[img][(size_t)y * L.w + x] = (uint8_t)c;
        }
}

void fill_colour(Level& L, int img) {
    L.col[img].resize((size_t)L.w * L.h);
    for (size_t i = 0; i < L.col[img].size(); i++) {
        const U4& p = L.rgba[img][i];
        L.col[img][i] = F3{(float)p.x / 255.f, (float)p.y / 255.f, (float)p.z / 255.f};
    }
}

// _d_bao_bilinear_resize<uchar4> (basic/bao_basic_cuda.cuh:565-601), generic ratio
U4 resize_at(const std::vector<U4>& src, int w, int h, int x, int y, float ratio) {
    const float div_scale = 1.f / ratio;
    const float fx = fmaf((float)(x + 1), div_scale, -1.f), fy = fmaf((float)(y + 1), div_scale, -1.f);
    const int xx = (int)fx, yy = (int)fy;
    const float dx = fmaxf(fminf(fx - (float)xx, 1.f), 0.f), dy = fmaxf(fminf(fy - (float)yy, 1.f), 0.f);
    float rx = 0.f, ry = 0.f, rz = 0.f;
    for (int m = 0; m <= 1; m++)
        for (int n = 0; n <= 1; n++) {
            const U4& p = src[(size_t)clampi(yy + n, 0, h - 1) * w + clampi(xx + m, 0, w - 1)];
            const float s = fabsf((float)(1 - m) - dx) * fabsf((float)(1 - n) - dy);
            rx = fmaf((float)p.x, s, rx);
            ry = fmaf((float)p.y, s, ry);
            rz = fmaf((float)p.z, s, rz);
        }
    return U4{(uint8_t)(unsigned)rx, (uint8_t)(unsigned)ry, (uint8_t)(unsigned)rz, 0};
}

// baoCudaPatchMatchMultiscalePrepare (bao_pmflow_refine_kernel.cu:1060-1071) + bao_cuda_construct_gauss_pyramid_pitched
// (basic/bao_basic_cuda.cuh:642-664).  As compiled, `int n = log(0.25)/log(ratio)` is 1 (double / float-log of 0.5f), so
// level 1 = resize(blur_{sigma 1, r 3}(level 0), 0.5) and level i>1 = resize(blur_{sigma 1, r 3}(level i-1),
// (float)pow(.5,i)*W0/W[i-1]).
void prepare(golden_ctx* c, const uint8_t* rgb1, const uint8_t* rgb2) {
    const uint8_t* rgb[2] = {rgb1, rgb2};
    Level& L0 = c->lv[0];
    for (int img = 0; img < 2; img++) {
        std::vector<U4> raw((size_t)L0.w * L0.h);  // bao_rgb2rgba (basic/bao_basic_cuda.h:258-267), alpha = 0
        for (size_t i = 0; i < raw.size(); i++) raw[i] = U4{rgb[img][3 * i], rgb[img][3 * i + 1], rgb[img][3 * i + 2], 0};
        L0.rgba[img].resize(raw.size());
        for (int y = 0; y < L0.h; y++)
            for (int x = 0; x < L0.w; x++) L0.rgba[img][(size_t)y * L0.w + x] = blur_at(raw, L0.w, L0.h, x, y, 2, kGaussR2Bits);  // :1063 sigma .5, r 2
        for (int i = 1; i < c->n_levels; i++) {
            Level& L = c->lv[i];
            const Level& S = c->lv[i - 1];
            const float ratio = i == 1 ? 0.5f : (float)pow((double)0.5f, i) * L0.w / S.w;
            L.rgba[img].resize((size_t)L.w * L.h);
            if (ratio == 0.5f) {
                // integral fx: the resize keeps exactly texel (2x+1, 2y+1) of the blurred source (weights 1,0,0,0)
                for (int y = 0; y < L.h; y++)
                    for (int x = 0; x < L.w; x++)
                        L.rgba[img][(size_t)y * L.w + x] = blur_at(S.rgba[img], S.w, S.h, clampi(2 * (x + 1) - 1, 0, S.w - 1), clampi(2 * (y + 1) - 1, 0, S.h - 1), 3, kGaussR3Bits);
            } else {
                std::vector<U4> blurred((size_t)S.w * S.h);
                for (int y = 0; y < S.h; y++)
                    for (int x = 0; x < S.w; x++) blurred[(size_t)y * S.w + x] = blur_at(S.rgba[img], S.w, S.h, x, y, 3, kGaussR3Bits);
                for (int y = 0; y < L.h; y++)
                    for (int x = 0; x < L.w; x++) L.rgba[img][(size_t)y * L.w + x] = resize_at(blurred, S.w, S.h, x, y, ratio);
            }
        }
        for (int i = 0; i < c->n_levels; i++) {
            fill_colour(c->lv[i], img);
            census_level(c->lv[i], img);  // :1067-1070
        }
    }
}

// ------------------------------------------------------------------------------------------------ XORWOW
// cuRAND XORWOW as the reference uses it (curand_init(1234, block_id, 0) bao_pmflow_kernel.cu:68; curand() :94-95).
// Published algorithm (Marsaglia xorwow + Weyl sequence, cuRAND device API): 160-bit xorshift state v[0..4], counter d;
//   t = v0 ^ (v0 >> 2); v0..v3 = v1..v4; v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1)); d += 362437; return v4 + d.
// A sub-sequence is 2^67 steps ahead of the previous one; the xorshift part is linear over GF(2), so the jump is a
// 160x160 bit-matrix power (67 squarings) applied once per set bit of the sub-sequence number; d is unchanged because
// 2^67 * k * 362437 = 0 mod 2^32.  Seeding constants as in curand_init.
struct Xorwow { uint32_t v[5], d; };
typedef std::vector<uint32_t> Mat;  // 160 rows x 5 words: row i = image of basis bit i

void xw_step_v(uint32_t* v) {
    const uint32_t t = v[0] ^ (v[0] >> 2);
    v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
    v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
}
void mat_apply(const Mat& m, const uint32_t* in, uint32_t* out) {
    uint32_t r[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 160; i++)
        if (in[i >> 5] >> (i & 31) & 1u)
            for (int k = 0; k < 5; k++) r[k] ^= m[i * 5 + k];
    memcpy(out, r, sizeof r);
}
Mat mat_mul(const Mat& a, const Mat& b) {  // (a then b)
    Mat c(160 * 5);
    for (int i = 0; i < 160; i++) mat_apply(b, &a[i * 5], &c[i * 5]);
    return c;
}
const std::vector<Mat>& jump_powers() {  // [j] = step^(2^67 * 2^j)
    static std::vector<Mat> pw;
    if (pw.empty()) {
        Mat m(160 * 5, 0);
        for (int i = 0; i < 160; i++) {
            uint32_t v[5] = {0, 0, 0, 0, 0};
            v[i >> 5] = 1u << (i & 31);
            xw_step_v(v);
            memcpy(&m[i * 5], v, sizeof v);
        }
        for (int s = 0; s < 67; s++) m = mat_mul(m, m);
        for (int j = 0; j < 32; j++) { pw.push_back(m); m = mat_mul(m, m); }
    }
    return pw;
}
Xorwow xw_init(unsigned long long seed, unsigned long long subseq) {
    Xorwow s;
    const uint32_t s0 = (uint32_t)seed ^ 0xaad26b49u, s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    s.d = 6615241u + t1 + t0;
    s.v[0] = 123456789u + t0; s.v[1] = 362436069u ^ t0; s.v[2] = 521288629u + t1; s.v[3] = 88675123u ^ t1; s.v[4] = 5783321u + t0;
    const std::vector<Mat>& pw = jump_powers();
    for (int j = 0; j < 32; j++)
        if (subseq >> j & 1ull) mat_apply(pw[j], s.v, s.v);
    return s;
}
inline uint32_t xw_next(Xorwow& s) {
    xw_step_v(s.v);
    s.d += 362437u;
    return s.v[4] + s.d;
}

// Random tables: d_setup_randgen + d_gen_rand_field (bao_pmflow_kernel.cu:50-109) and the draws of d_update_random_guess
// (:1537-1551); one stream per 16x16 block id, same for both directions and every pair.
// patch scale of the scaled PatchMatch from the SECOND draw of a pixel: float((10 + ((rdn2 % PM_SCALE_RANGE) - PM_SCALE_MIN)) / float(10.0f))
// in unsigned arithmetic (bao_pmflow_kernel.cu:138, :1631; defs.h:40-41: 9 and 4) = (r % 9 + 6) / 10 in [0.6, 1.4]
inline float scale_of_draw(uint32_t r2) { return (float)(r2 % 9u + 6u) / 10.0f; }

void build_rng(golden_ctx* c) {
    const Level& L = c->lv[c->n_levels - 1];
    const int gx = (L.w + 15) / 16, gy = (L.h + 15) / 16;
    const size_t n = (size_t)L.w * L.h;
    c->rng_init.assign(n, S2{0, 0});
    c->rng_search.assign(n * c->num_iter * c->num_guess, S2{0, 0});
    c->rng_init_scale.assign(n, 0.f);
    c->rng_search_scale.assign(n * c->num_iter * c->num_guess, 0.f);
    for (int bid = 0; bid < gx * gy; bid++) {
        Xorwow st = xw_init(1234ull, (unsigned long long)bid);
        const int bx = bid % gx, by = bid / gx;
        for (int i = 0; i < 16; i++)
            for (int j = 0; j < 16; j++) {
                const uint32_t r1 = xw_next(st), r2 = xw_next(st);
                const int x = bx * 16 + j, y = by * 16 + i;
                if (x < L.w && y < L.h) {
                    c->rng_init[(size_t)y * L.w + x] = S2{(int16_t)(r1 % (uint32_t)(L.w + 1)), (int16_t)(r2 % (uint32_t)(L.h + 1))};
                    c->rng_init_scale[(size_t)y * L.w + x] = scale_of_draw(r2);
                }
            }
        for (int it = 0; it < c->num_iter; it++)
            for (int k = 0; k < c->num_guess; k++)
                for (int i = 0; i < 16; i++)
                    for (int j = 0; j < 16; j++) {
                        const uint32_t r1 = xw_next(st), r2 = xw_next(st);
                        const int x = bx * 16 + j, y = by * 16 + i;
                        if (x < L.w && y < L.h) {
                            c->rng_search[((size_t)(it * c->num_guess + k) * L.h + y) * L.w + x] = S2{(int16_t)r1, (int16_t)r2};
                            c->rng_search_scale[((size_t)(it * c->num_guess + k) * L.h + y) * L.w + x] = scale_of_draw(r2);
                        }
                    }
    }
}

// ------------------------------------------------------------------------------------------------ patch cost
// One sample of _d_compute_patch_dist (bao_pmflow_kernel.cu:274-296) in the operation order nvcc gives it:
//   c = max|p1-p2|; cost = (1 - exp(c*c / -0.01)) + LUT[popc];  arg = fma(d1,d1, d2*d2); w = exp(arg / -0.01) * (G|j|*G|i|);
//   cost_sum = fma(cost, w, cost_sum); weight_sum += w.
inline void sample(const golden_ctx* c, const F3& p1, uint8_t cen1, const F3& p2, uint8_t cen2, const F3& c1, const F3& c2, int ai, int aj,
                   float& cs, float& ws) {
    const float neg = -(0.1f * 0.1f);  // LAMBDA_AD^2 = PM_SIG_R^2, folded in float by the compiler
    const float cc = max3abs(p1, p2);
    const float e = expf_dev((cc * cc) / neg);
    const float cost = (1.0f - e) + c->census_lut[__builtin_popcount((unsigned)(cen1 ^ cen2))];
    const float d1 = max3abs(c1, p1), d2 = max3abs(c2, p2);
    const float arg = fmaf(d1, d1, d2 * d2);
    const float w = expf_dev(arg / neg) * (c->G[aj] * c->G[ai]);
    cs = fmaf(cost, w, cs);
    ws = ws + w;
}

// _d_compute_patch_dist (:255-301): 10x10 samples at stride 2, texture clamp addressing on both images
float patch_cost(const golden_ctx* c, const Level& L, int A, int B, int x1, int y1, int x2, int y2) {
    const F3 c1 = L.C(A, x1, y1), c2 = L.C(B, x2, y2);
    float cs = 0.f, ws = 0.f;
    for (int i = -PATCH_R; i <= PATCH_R; i += c->stride)
        for (int j = -PATCH_R; j <= PATCH_R; j += c->stride)
            sample(c, L.C(A, x1 + j, y1 + i), L.Cen(A, x1 + j, y1 + i), L.C(B, x2 + j, y2 + i), L.Cen(B, x2 + j, y2 + i), c1, c2, abs(i), abs(j), cs, ws);
    return cs / ws;
}

// _d_compute_patch_dist_planefitting (:334-513): min over the identity and three affine sample maps; sample coordinates
// cx2 = fma(i, C_uy, fma(j, C_ux, float(x1+j) + float(x2-x1))) (cy2 alike), point-sampled = floor + clamp (probed).
const float kPF[3][4] = {{0.177f, -0.011f, -0.003f, 0.301f}, {0.125f, -0.357f, 0.009f, 0.308f}, {0.205f, 0.370f, 0.011f, 0.296f}};  // :319-332
float patch_cost_pf(const golden_ctx* c, const Level& L, int x1, int y1, int x2, int y2) {
    const F3 c1 = L.C(0, x1, y1), c2 = L.C(1, x2, y2);
    const float uu = (float)(x2 - x1), vv = (float)(y2 - y1);
    float best = 0.f;
    for (int q = 0; q < 4; q++) {
        float cs = 0.f, ws = 0.f;
        for (int i = -PATCH_R; i <= PATCH_R; i += c->stride)
            for (int j = -PATCH_R; j <= PATCH_R; j += c->stride) {
                float cx2 = (float)(x1 + j) + uu, cy2 = (float)(y1 + i) + vv;
                if (q > 0) {
                    cx2 = fmaf((float)i, kPF[q - 1][1], fmaf((float)j, kPF[q - 1][0], cx2));
                    cy2 = fmaf((float)i, kPF[q - 1][3], fmaf((float)j, kPF[q - 1][2], cy2));
                }
                const int sx = (int)floorf(cx2), sy = (int)floorf(cy2);
                sample(c, L.C(0, x1 + j, y1 + i), L.Cen(0, x1 + j, y1 + i), L.C(1, sx, sy), L.Cen(1, sx, sy), c1, c2, abs(i), abs(j), cs, ws);
            }
        const float k = cs / ws;
        best = q == 0 ? k : (k < best ? k : best);  // :512 nested __min; evaluated innermost-first, same winner
    }
    return best;
}

// cost of a PatchMatch candidate: the plain patch distance, or -- forward direction only -- the plane-fitting one
inline float pm_cost(const golden_ctx* c, const Level& L, int A, int B, int x1, int y1, int x2, int y2) {
    return (c->pf_cost && A == 0) ? patch_cost_pf(c, L, x1, y1, x2, y2) : patch_cost(c, L, A, B, x1, y1, x2, y2);
}

// ------------------------------------------------------------------------------------------------ PatchMatch
// baoCudaPatchMatch (:1760-1826) for one direction.  Segment passes run in lock-step (step t of every segment before
// step t+1), the order the reference gets from warp-synchronous execution (DESIGN.md "racy stages").
void pm_propagate(golden_ctx* c, int dir, int pass) {
    const Level& L = c->lv[c->n_levels - 1];
    const int A = dir, B = dir ^ 1;
    const bool row = (pass == 0 || pass == 2), fwd = pass < 2;
    const int n_line = row ? L.h : L.w, len = row ? L.w : L.h, sl = c->seg_len;
    const int n_seg = (len + sl - 1) / sl;
    std::vector<S2>& nnf = c->nnf[dir];
    std::vector<float>& cost = c->cost[dir];
    std::vector<S2> prev((size_t)n_line * n_seg);
    std::vector<int> start((size_t)n_seg), steps((size_t)n_seg);
    for (int s = 0; s < n_seg; s++) {
        if (fwd) {  // :1055-1058
            start[s] = s == 0 ? 0 : s * sl - 1;
            steps[s] = (len - 1 < start[s] + sl ? len - 1 : start[s] + sl) - start[s];
        } else {    // :1085-1088
            start[s] = (s + 1) * sl >= len ? len - 1 : (s + 1) * sl;
            steps[s] = start[s] - s * sl;
        }
    }
    auto idx = [&](int line, int i) -> size_t { return row ? (size_t)line * L.w + i : (size_t)i * L.w + line; };
    for (int line = 0; line < n_line; line++)
        for (int s = 0; s < n_seg; s++) prev[(size_t)line * n_seg + s] = nnf[idx(line, start[s])];
    for (int t = 1; t <= sl; t++)
        for (int line = 0; line < n_line; line++)
            for (int s = 0; s < n_seg; s++) {
                if (t > steps[s]) continue;
                const int i = fwd ? start[s] + t : start[s] - t;
                S2& p = prev[(size_t)line * n_seg + s];
                if (pass == 0) p.x = (int16_t)(p.x + 1 < L.w - 1 ? p.x + 1 : L.w - 1);  // :1065
                if (pass == 1) p.y = (int16_t)(p.y + 1 < L.h - 1 ? p.y + 1 : L.h - 1);  // :1125
                if (pass == 2) p.x = (int16_t)(p.x - 1 > 0 ? p.x - 1 : 0);              // :1095
                if (pass == 3) p.y = (int16_t)(p.y - 1 > 0 ? p.y - 1 : 0);              // :1155
                const size_t id = idx(line, i);
                const float cv = pm_cost(c, L, A, B, row ? i : line, row ? line : i, p.x, p.y);
                if (cv < cost[id]) { nnf[id] = p; cost[id] = cv; } else { p = nnf[id]; }
            }
}

void pm_search(golden_ctx* c, int dir, int it) {  // d_update_random_guess (:1519-1586)
    const Level& L = c->lv[c->n_levels - 1];
    const int A = dir, B = dir ^ 1;
    const size_t n = (size_t)L.w * L.h;
    for (int y = 0; y < L.h; y++)
        for (int x = 0; x < L.w; x++) {
            const size_t id = (size_t)y * L.w + x;
            S2 best = c->nnf[dir][id];
            const S2 entry = best;
            float best_cost = c->cost[dir][id];
            int mag = c->search_range;
            for (int k = 0; k < c->num_guess; k++) {
                const S2 rr = c->rng_search[(size_t)(it * c->num_guess + k) * n + id];
                const uint32_t r1 = (uint32_t)(int32_t)rr.x, r2 = (uint32_t)(int32_t)rr.y;  // short -> unsigned (:1557-1558)
                const int16_t xmin = (int16_t)(entry.x - mag > 0 ? entry.x - mag : 0), xmax = (int16_t)(entry.x + mag + 1 < L.w + 1 ? entry.x + mag + 1 : L.w + 1);
                const int16_t ymin = (int16_t)(entry.y - mag > 0 ? entry.y - mag : 0), ymax = (int16_t)(entry.y + mag + 1 < L.h + 1 ? entry.y + mag + 1 : L.h + 1);
                const int16_t gx = (int16_t)(xmin + r1 % (uint32_t)(xmax - xmin)), gy = (int16_t)(ymin + r2 % (uint32_t)(ymax - ymin));
                if (mag / 2 >= c->radius_min) mag /= 2;
                const float cv = pm_cost(c, L, A, B, x, y, gx, gy);
                if (cv < best_cost) { best = S2{gx, gy}; best_cost = cv; }
            }
            c->nnf[dir][id] = best;
            c->cost[dir][id] = best_cost;
        }
}

void patchmatch(golden_ctx* c, int n_steps) {
    const Level& L = c->lv[c->n_levels - 1];
    const size_t n = (size_t)L.w * L.h;
    if (n_steps < 0) n_steps = 1 << 30;
    for (int dir = 0; dir < 2; dir++) {
        int step = 0;
        if (step++ >= n_steps) continue;
        c->nnf[dir] = c->rng_init;  // baoGenerateRandomField
        c->cost[dir].resize(n);
        for (int y = 0; y < L.h; y++)  // baoComputeCostField (:636-645)
            for (int x = 0; x < L.w; x++) {
                const S2 t = c->nnf[dir][(size_t)y * L.w + x];
                c->cost[dir][(size_t)y * L.w + x] = pm_cost(c, L, dir, dir ^ 1, x, y, t.x, t.y);
            }
        for (int it = 0; it < c->num_iter; it++) {
            bool stop = false;
            for (int pass = 0; pass < 4 && !stop; pass++) {
                if (step++ >= n_steps) { stop = true; break; }
                pm_propagate(c, dir, pass);
            }
            if (stop || step++ >= n_steps) break;
            pm_search(c, dir, it);
        }
    }
}

// ------------------------------------------------------------------------------------------------ consistency
// ------------------------------------------------------------------------------------------------ scaled PatchMatch
// baoCudaPatchMatch_Scaled (bao_pmflow_kernel.cu:1828-1895): forward direction over (target, patch scale).  Unfinished upstream and restated AS
// IT STANDS: _d_compute_patch_dist_scaled (:588-634) is the bilateral AD term alone (census lines commented out), image-2 samples at
// (x2 + float(j)*scale, y2 + float(i)*scale) -- contracted by nvcc to fma(scale, float(j), float(x2)) -- through a point-filtered clamped
// fetch; d_row_propagate_seg_scaled stores the winning candidate's SCALE into the cost plane (:1207); nothing may be skipped.
float patch_cost_scaled(const golden_ctx* c, const Level& L, int x1, int y1, int x2, int y2, float scale) {
    const F3 c1 = L.C(0, x1, y1), c2 = L.C(1, x2, y2);
    const float neg = -(0.1f * 0.1f);
    float cs = 0.f, ws = 0.f;
    for (int i = -PATCH_R; i <= PATCH_R; i += 2)
        for (int j = -PATCH_R; j <= PATCH_R; j += 2) {
            const int sx = (int)floorf(fmaf(scale, (float)j, (float)x2)), sy = (int)floorf(fmaf(scale, (float)i, (float)y2));
            const F3 p1 = L.C(0, x1 + j, y1 + i), p2 = L.C(1, sx, sy);
            const float cc = max3abs(p1, p2);
            const float cost = 1.0f - expf_dev((cc * cc) / neg);                                   // :610-611
            const float d1 = max3abs(c1, p1), d2 = max3abs(c2, p2);
            const float w = expf_dev(fmaf(d1, d1, d2 * d2) / neg) * (c->G[abs(j)] * c->G[abs(i)]);  // :613-618
            cs = fmaf(cost, w, cs);
            ws = ws + w;
        }
    return cs / ws;
}

void patchmatch_scaled(golden_ctx* c) {
    const Level& L = c->lv[c->n_levels - 1];
    const size_t n = (size_t)L.w * L.h;
    std::vector<S2>& nnf = c->nnf[0];
    std::vector<float>& cost = c->cost[0];
    std::vector<float>& scl = c->scale;
    nnf = c->rng_init;            // baoGenerateRandomField_Scaled (:167-178)
    scl = c->rng_init_scale;
    cost.resize(n);
    for (int y = 0; y < L.h; y++)  // baoComputeCostField_Scaled (:647-656)
        for (int x = 0; x < L.w; x++) {
            const size_t id = (size_t)y * L.w + x;
            cost[id] = patch_cost_scaled(c, L, x, y, nnf[id].x, nnf[id].y, scl[id]);
        }
    const int sl = c->seg_len;
    for (int it = 0; it < c->num_iter; it++) {
        for (int pass = 0; pass < 4; pass++) {   // baoSegPropagate_Scaled (:1318-1331): row, column, row reverse, column reverse; lock-step like pm_propagate
            const bool row = (pass == 0 || pass == 2), fwd = pass < 2;
            const int n_line = row ? L.h : L.w, len = row ? L.w : L.h;
            const int n_seg = (len + sl - 1) / sl;
            std::vector<S2> prev((size_t)n_line * n_seg);
            std::vector<float> prev_s((size_t)n_line * n_seg);
            std::vector<int> start((size_t)n_seg), steps((size_t)n_seg);
            for (int s = 0; s < n_seg; s++) {
                if (fwd) { start[s] = s == 0 ? 0 : s * sl - 1; steps[s] = (len - 1 < start[s] + sl ? len - 1 : start[s] + sl) - start[s]; }   // :1189-1192
                else { start[s] = (s + 1) * sl >= len ? len - 1 : (s + 1) * sl; steps[s] = start[s] - s * sl; }                               // :1225-1227
            }
            auto idx = [&](int line, int i) -> size_t { return row ? (size_t)line * L.w + i : (size_t)i * L.w + line; };
            for (int line = 0; line < n_line; line++)
                for (int s = 0; s < n_seg; s++) {
                    prev[(size_t)line * n_seg + s] = nnf[idx(line, start[s])];
                    prev_s[(size_t)line * n_seg + s] = scl[idx(line, start[s])];
                }
            for (int t = 1; t <= sl; t++)
                for (int line = 0; line < n_line; line++)
                    for (int s = 0; s < n_seg; s++) {
                        if (t > steps[s]) continue;
                        const int i = fwd ? start[s] + t : start[s] - t;
                        S2& p = prev[(size_t)line * n_seg + s];
                        float& ps = prev_s[(size_t)line * n_seg + s];
                        if (pass == 0) p.x = (int16_t)(p.x + 1 < L.w - 1 ? p.x + 1 : L.w - 1);
                        if (pass == 1) p.y = (int16_t)(p.y + 1 < L.h - 1 ? p.y + 1 : L.h - 1);
                        if (pass == 2) p.x = (int16_t)(p.x - 1 > 0 ? p.x - 1 : 0);
                        if (pass == 3) p.y = (int16_t)(p.y - 1 > 0 ? p.y - 1 : 0);
                        const size_t id = idx(line, i);
                        const float cv = patch_cost_scaled(c, L, row ? i : line, row ? line : i, p.x, p.y, ps);
                        if (cv < cost[id]) { nnf[id] = p; scl[id] = ps; cost[id] = pass == 0 ? ps : cv; }   // :1207: the forward row pass stores the scale
                        else { p = nnf[id]; ps = scl[id]; }
                    }
        }
        for (int y = 0; y < L.h; y++)   // d_update_random_guess_scaled (:1596-1671)
            for (int x = 0; x < L.w; x++) {
                const size_t id = (size_t)y * L.w + x;
                S2 best = nnf[id];
                const S2 entry = best;
                float best_s = scl[id], best_cost = cost[id];
                int mag = c->search_range;
                for (int k = 0; k < c->num_guess; k++) {
                    const S2 rr = c->rng_search[(size_t)(it * c->num_guess + k) * n + id];
                    const float gs = c->rng_search_scale[(size_t)(it * c->num_guess + k) * n + id];
                    const uint32_t r1 = (uint32_t)(int32_t)rr.x, r2 = (uint32_t)(int32_t)rr.y;
                    const int16_t xmin = (int16_t)(entry.x - mag > 0 ? entry.x - mag : 0), xmax = (int16_t)(entry.x + mag + 1 < L.w + 1 ? entry.x + mag + 1 : L.w + 1);
                    const int16_t ymin = (int16_t)(entry.y - mag > 0 ? entry.y - mag : 0), ymax = (int16_t)(entry.y + mag + 1 < L.h + 1 ? entry.y + mag + 1 : L.h + 1);
                    const int16_t gx = (int16_t)(xmin + r1 % (uint32_t)(xmax - xmin)), gy = (int16_t)(ymin + r2 % (uint32_t)(ymax - ymin));
                    if (mag / 2 >= c->radius_min) mag /= 2;
                    const float cv = patch_cost_scaled(c, L, x, y, gx, gy, gs);
                    if (cv < best_cost) { best = S2{gx, gy}; best_s = gs; best_cost = cv; }
                }
                nnf[id] = best; scl[id] = best_s; cost[id] = best_cost;
            }
    }
}

void lr_check(golden_ctx* c, int dir) {  // d_left_right_check (bao_pmflow_refine_kernel.cu:53-76)
    const Level& L = c->lv[c->n_levels - 1];
    for (int y = 0; y < L.h; y++)
        for (int x = 0; x < L.w; x++) {
            const size_t id = (size_t)y * L.w + x;
            const S2 d = c->nnf[dir][id];
            bool bad = d.y < 0 || d.y >= L.h || d.x < 0 || d.x >= L.w;
            if (!bad) {
                const S2 d2 = c->nnf[dir ^ 1][(size_t)d.y * L.w + d.x];
                bad = abs(d2.x - x) > 0 || abs(d2.y - y) > 0;
            }
            if (bad) { c->nnf[dir][id] = S2{(int16_t)INVALID_LOCATION, (int16_t)INVALID_LOCATION}; c->cost[dir][id] = FLT_MAX; }
        }
}

void outlier_removal(golden_ctx* c) {  // d_outlier_removal (:149-182), snapshot semantics
    const Level& L = c->lv[c->n_levels - 1];
    const int R = 6, sim = 2, thresh = (2 * R + 1) * (2 * R + 1) / 2;  // defs.h:68, :146-147
    const std::vector<S2> src = c->nnf[0];
    for (int y = 0; y < L.h; y++)
        for (int x = 0; x < L.w; x++) {
            const S2 cur = src[(size_t)y * L.w + x];
            if (cur.x < 0 && cur.y < 0) continue;
            const int ux = cur.x - x, uy = cur.y - y;
            int count = 0;
            for (int dy = -R; dy <= R; dy++)
                for (int dx = -R; dx <= R; dx++) {
                    const int cx = x + dx, cy = y + dy;
                    if (cx < 0 || cy < 0 || cx >= L.w || cy >= L.h) continue;
                    const S2 nb = src[(size_t)cy * L.w + cx];
                    if (abs((int16_t)(nb.x - cx) - ux) <= sim && abs((int16_t)(nb.y - cy) - uy) <= sim) count++;
                }
            if (count < thresh) {
                c->nnf[0][(size_t)y * L.w + x] = S2{(int16_t)INVALID_LOCATION, (int16_t)INVALID_LOCATION};
                c->cost[0][(size_t)y * L.w + x] = FLT_MAX;
            }
        }
}

void wmf_sweep(golden_ctx* c, bool only_occlusion) {  // d_weighted_median_filtering (:206-259), snapshot semantics
    const Level& L = c->lv[c->n_levels - 1];
    const int R = 4;
    const float neg = -(0.02f * 0.02f);  // WMF_SIG_R^2 (defs.h:60)
    const std::vector<S2> src = c->nnf[0];
    for (int y = 0; y < L.h; y++)
        for (int x = 0; x < L.w; x++) {
            S2 out = src[(size_t)y * L.w + x];
            if (only_occlusion && out.x >= 0 && out.y >= 0) continue;
            const F3 cc = L.C(0, x, y);
            float wgt[81]; int du[81], dv[81]; bool ok[81];
            int k = 0;
            for (int dy = -R; dy <= R; dy++)
                for (int dx = -R; dx <= R; dx++, k++) {
                    const int cx = x + dx, cy = y + dy;
                    ok[k] = false;
                    if (cx < 0 || cy < 0 || cx >= L.w || cy >= L.h) continue;
                    const S2 t = src[(size_t)cy * L.w + cx];
                    if (t.x < 0 || t.y < 0) continue;
                    ok[k] = true; du[k] = t.x - cx; dv[k] = t.y - cy;
                    const float dr = max3abs(L.C(0, cx, cy), cc);
                    wgt[k] = expf_dev((dr * dr) / neg) * (c->wmf_g[abs(dx)] * c->wmf_g[abs(dy)]);  // :198-204
                }
            float best = FLT_MAX;
            for (int a = 0; a < 81; a++) {
                if (!ok[a]) continue;
                float cs = 0.f, ws = 0.f;
                for (int q = 0; q < 81; q++) {
                    if (!ok[q]) continue;
                    const int dist = abs(du[a] - du[q]) > abs(dv[a] - dv[q]) ? abs(du[a] - du[q]) : abs(dv[a] - dv[q]);
                    cs = fmaf(wgt[q], (float)dist, cs);  // :244
                    ws = ws + wgt[q];
                }
                if (ws > 0.0f && cs < best) { best = cs; out = S2{(int16_t)(du[a] + x), (int16_t)(dv[a] + y)}; }
            }
            if (out.x < 0 || out.y < 0) continue;  // :257
            c->nnf[0][(size_t)y * L.w + x] = out;
        }
}

void fill_holes(golden_ctx* c) {  // d_fill_holes (:297-371), snapshot semantics
    const Level& L = c->lv[c->n_levels - 1];
    const std::vector<S2> src = c->nnf[0];
    auto F = [&](int x, int y) { return src[(size_t)y * L.w + x]; };
    for (int y = 0; y < L.h; y++)
        for (int x = 0; x < L.w; x++) {
            S2 cur = F(x, y);
            if (cur.x >= 0 && cur.y >= 0) continue;
            S2 nd[4] = {cur, cur, cur, cur};
            int nx[4] = {x, x, x, x}, ny[4] = {y, y, y, y};
            for (int cx = x - 1; cx >= 0; cx--) { nd[0] = F(cx, y); if (nd[0].x >= 0 && nd[0].y >= 0) { nx[0] = cx; break; } }
            for (int cx = x + 1; cx < L.w; cx++) { nd[1] = F(cx, y); if (nd[1].x >= 0 && nd[1].y >= 0) { nx[1] = cx; break; } }
            for (int cy = y - 1; cy >= 0; cy--) { nd[2] = F(x, cy); if (nd[2].x >= 0 && nd[2].y >= 0) { ny[2] = cy; break; } }
            for (int cy = y + 1; cy < L.h; cy++) { nd[3] = F(x, cy); if (nd[3].x >= 0 && nd[3].y >= 0) { ny[3] = cy; break; } }
            const F3 cc = L.C(0, x, y);
            float md = FLT_MAX;
            for (int i = 0; i < 4; i++) {
                const float diff = max3abs(L.C(0, nx[i], ny[i]), cc);
                if (diff < md && nd[i].x >= 0 && nd[i].y >= 0) { md = diff; cur.x = (int16_t)(nd[i].x - nx[i]); cur.y = (int16_t)(nd[i].y - ny[i]); }
            }
            cur.x = (int16_t)(cur.x + x); cur.y = (int16_t)(cur.y + y);  // :368-370 (also when nothing was found)
            c->nnf[0][(size_t)y * L.w + x] = cur;
        }
}

void consistency(golden_ctx* c) {  // bao_flow_patchmatch_multiscale_cuda.cpp:233-258
    const int Lc = c->n_levels - 1;
    Level& L = c->lv[Lc];
    lr_check(c, 0);
    lr_check(c, 1);
    outlier_removal(c);
    for (int i = 0; i < 20; i++) wmf_sweep(c, true);
    fill_holes(c);
    L.flow.resize((size_t)L.w * L.h);
    for (int y = 0; y < L.h; y++)  // d_convert_nnf_to_flow (:636-655)
        for (int x = 0; x < L.w; x++) {
            const S2 d = c->nnf[0][(size_t)y * L.w + x];
            L.flow[(size_t)y * L.w + x] = (d.x <= INVALID_LOCATION || d.y <= INVALID_LOCATION) ? F2{UNKNOWN_FLOW, UNKNOWN_FLOW} : F2{(float)(d.x - x), (float)(d.y - y)};
        }
}

// ------------------------------------------------------------------------------------------------ coarse to fine
F2 upsample2(const Level& S, int x, int y) {  // _d_bao_bilinear_resize<float2> ratio 2, then x2.0 (basic/bao_basic_cuda.cuh:511-537,135-149)
    const float fx = fmaf((float)(x + 1), 0.5f, -1.f), fy = fmaf((float)(y + 1), 0.5f, -1.f);
    const int xx = (int)fx, yy = (int)fy;
    const float dx = fmaxf(fminf(fx - (float)xx, 1.f), 0.f), dy = fmaxf(fminf(fy - (float)yy, 1.f), 0.f);
    float rx = 0.f, ry = 0.f;
    for (int m = 0; m <= 1; m++)
        for (int n = 0; n <= 1; n++) {
            const F2& p = S.flow[(size_t)clampi(yy + n, 0, S.h - 1) * S.w + clampi(xx + m, 0, S.w - 1)];
            const float s = fabsf((float)(1 - m) - dx) * fabsf((float)(1 - n) - dy);
            rx = fmaf(s, p.x, rx);
            ry = fmaf(s, p.y, ry);
        }
    return F2{rx * 2.0f, ry * 2.0f};
}

void refine_level(golden_ctx* c, int l) {  // d_bilateral_refine_flow_planefitting (bao_pmflow_kernel.cu:2005-2041)
    Level& L = c->lv[l];
    const Level& S = c->lv[l + 1];
    L.flow.resize((size_t)L.w * L.h);
    for (int y = 0; y < L.h; y++)
        for (int x = 0; x < L.w; x++) {
            const F2 fl = upsample2(S, x, y);
            F2 out;
            if (fl.x > UNKNOWN_FLOW_THRESH || fl.y > UNKNOWN_FLOW_THRESH) {
                out = F2{0.f, 0.f};
            } else {
                const int16_t cxc = (int16_t)((int16_t)(int)fl.x + x), cyc = (int16_t)((int16_t)(int)fl.y + y);
                int16_t bx = cxc, by = cyc;
                float best = 999999.f;
                for (int m = 0; m < 3; m++)
                    for (int n = 0; n < 3; n++) {
                        const int16_t cx = (int16_t)(cxc + m - 1), cy = (int16_t)(cyc + n - 1);
                        if (cx < 0 || cy < 0 || cx >= L.w || cy >= L.h) continue;
                        const float cv = patch_cost_pf(c, L, x, y, cx, cy);
                        if (cv < best) { best = cv; bx = cx; by = cy; }
                    }
                out = F2{(float)(bx - x), (float)(by - y)};
            }
            L.flow[(size_t)y * L.w + x] = out;
        }
}

void smooth_level(golden_ctx* c, int l) {  // d_flow_bilateral_filtering (bao_pmflow_refine_kernel.cu:764-799), snapshot semantics
    Level& L = c->lv[l];
    const int R = 10;
    const float neg = -(0.02f * 0.02f);
    const std::vector<F2> src = L.flow;
    for (int y = 0; y < L.h; y++)
        for (int x = 0; x < L.w; x++) {
            const F3 cc = L.C(0, x, y);
            float nx = 0.f, ny = 0.f, ws = 0.f;
            for (int dy = -R; dy <= R; dy++)
                for (int dx = -R; dx <= R; dx++) {
                    const int cx = x + dx, cy = y + dy;
                    if (cx < 0 || cy < 0 || cx >= L.w || cy >= L.h) continue;
                    const F2 f = src[(size_t)cy * L.w + cx];
                    if (f.x > UNKNOWN_FLOW_THRESH || f.y > UNKNOWN_FLOW_THRESH) continue;
                    const float dr = max3abs(L.C(0, cx, cy), cc);
                    const float wg = expf_dev((dr * dr) / neg) * (c->blf_g[abs(dx)] * c->blf_g[abs(dy)]);
                    nx = fmaf(wg, f.x, nx);
                    ny = fmaf(wg, f.y, ny);
                    ws = ws + wg;
                }
            if (ws != 0.f) L.flow[(size_t)y * L.w + x] = F2{nx / ws, ny / ws};
        }
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

void golden_level_dims_for(int h, int w, int level, int* oh, int* ow) { level_dims(h, w, level, oh, ow); }

golden_ctx* golden_create(int h, int w, int levels, int num_iter) {
    golden_ctx* c = new golden_ctx;
    c->n_levels = levels;
    c->num_iter = num_iter;
    c->lv.resize(levels);
    for (int i = 0; i < levels; i++) level_dims(h, w, i, &c->lv[i].h, &c->lv[i].w);
    // _initGaussianLookupTable (bao_pmflow_kernel.cu:670-687) and the filter LUTs (refine:270-275, :809-813), host expf
    volatile float sig_s = 0.5f * PATCH_R;
    for (int i = 0; i <= PATCH_R; i++) c->G[i] = expf(-(i * i) / (sig_s * sig_s));
    volatile float lc = 0.3f;
    for (int i = 0; i <= 8; i++) c->census_lut[i] = 1 - expf(-float(i * i) / (lc * 8 * lc * 8));
    volatile float ws = 4.0f;
    for (int i = 0; i <= 4; i++) c->wmf_g[i] = expf(-float(i * i) / (ws * ws));
    volatile int bs = 5;
    for (int i = 0; i <= 10; i++) c->blf_g[i] = expf(-float(i * i) / float(bs * bs));
    build_rng(c);
    return c;
}
void golden_destroy(golden_ctx* c) { delete c; }
void golden_set_stride(golden_ctx* c, int stride) { c->stride = stride; }
void golden_set_pf_cost(golden_ctx* c, int on) { c->pf_cost = on; }
int golden_num_levels(const golden_ctx* c) { return c->n_levels; }
void golden_level_dims(const golden_ctx* c, int level, int* h, int* w) { *h = c->lv[level].h; *w = c->lv[level].w; }

void golden_prepare(golden_ctx* c, const uint8_t* rgb1, const uint8_t* rgb2) { prepare(c, rgb1, rgb2); }
void golden_patchmatch(golden_ctx* c, int n_steps) { patchmatch(c, n_steps); }
void golden_patchmatch_scaled(golden_ctx* c) { patchmatch_scaled(c); }   // forward field in planes nnf_fwd / cost_fwd, scales in plane 9
void golden_consistency(golden_ctx* c) { consistency(c); }
void golden_c2f(golden_ctx* c, float* flow_uv) {
    for (int l = c->n_levels - 2; l >= 0; l--) {  // bao_flow_patchmatch_multiscale_cuda.cpp:275-282
        refine_level(c, l);
        smooth_level(c, l);
    }
    smooth_level(c, 0);  // :289
    if (flow_uv) memcpy(flow_uv, c->lv[0].flow.data(), c->lv[0].flow.size() * sizeof(F2));
}
void golden_compute(golden_ctx* c, const uint8_t* rgb1, const uint8_t* rgb2, float* flow_uv) {
    prepare(c, rgb1, rgb2);
    patchmatch(c, -1);
    consistency(c);
    golden_c2f(c, flow_uv);
}

long golden_read_plane(golden_ctx* c, int which, int level, void* out) {
    const Level& L = c->lv[level];
    const Level& Lc = c->lv[c->n_levels - 1];
    const size_t n = (size_t)L.w * L.h, nc = (size_t)Lc.w * Lc.h;
    switch (which) {
    case 0: case 1: if (L.rgba[which].size() != n) return -1; memcpy(out, L.rgba[which].data(), n * 4); return (long)n * 4;
    case 2: case 3: if (L.census[which - 2].size() != n) return -1; memcpy(out, L.census[which - 2].data(), n); return (long)n;
    case 4: case 5: if (c->nnf[which - 4].size() != nc) return -1; memcpy(out, c->nnf[which - 4].data(), nc * 4); return (long)nc * 4;
    case 6: case 7: if (c->cost[which - 6].size() != nc) return -1; memcpy(out, c->cost[which - 6].data(), nc * 4); return (long)nc * 4;
    case 8: if (L.flow.size() != n) return -1; memcpy(out, L.flow.data(), n * 8); return (long)n * 8;
    case 9: if (c->scale.size() != nc) return -1; memcpy(out, c->scale.data(), nc * 4); return (long)nc * 4;
    }
    return -1;
}
long golden_write_plane(golden_ctx* c, int which, int level, const void* in) {
    Level& L = c->lv[level];
    const Level& Lc = c->lv[c->n_levels - 1];
    const size_t n = (size_t)L.w * L.h, nc = (size_t)Lc.w * Lc.h;
    switch (which) {
    case 4: case 5: c->nnf[which - 4].resize(nc); memcpy(c->nnf[which - 4].data(), in, nc * 4); return (long)nc * 4;
    case 6: case 7: c->cost[which - 6].resize(nc); memcpy(c->cost[which - 6].data(), in, nc * 4); return (long)nc * 4;
    case 8: L.flow.resize(n); memcpy(L.flow.data(), in, n * 8); return (long)n * 8;
    }
    return -1;
}

void golden_xorwow(unsigned long long seed, unsigned long long subsequence, int n, uint32_t* out) {
    Xorwow s = xw_init(seed, subsequence);
    for (int i = 0; i < n; i++) out[i] = xw_next(s);
}

float golden_patch_cost(golden_ctx* c, int dir, int x1, int y1, int x2, int y2) {
    return patch_cost(c, c->lv[c->n_levels - 1], dir, dir ^ 1, x1, y1, x2, y2);
}

}  // extern "C"
