// TEST INFRASTRUCTURE ONLY — same scope and rules as golden.cpp (see golden.h): only tests/ may call this.
//
// Single-threaded CPU restatement of the reference's optional sub-pixel stage:
//   golden_census_bicubic   baoCudaCensusTransform_Bicubic  (bao_pmflow_census_kernel.cu:115-181, 3rdparty/nv-cuda-v5.0/bicubicTexture_kernel.cuh:33-104)
//   golden_subpix_refine    baoCudaSubpixRefine             (bao_pmflow_refine_kernel.cu:395-634, 678-722)
// Plain loops, one cost evaluation at a time exactly like the reference's kernel (no tables, no reuse), float arithmetic with
// fmaf() where nvcc contracts the reference's expressions (read from the PTX/SASS of the reference build), -ffp-contract=off.
//
// Parity status.  The census is pure IEEE arithmetic on point-sampled texels -> pinned BIT-EXACT by tests/golden/refsub_*.npz
// (outputs of the reference build on a B200).  The refinement depends on two hardware units a CPU can only approximate:
//   * the texture unit's bilinear filter: the reference binds the images with cudaFilterModeLinear and fetches at integer
//     coordinates, i.e. half way between texels; restated as the mean of the 2x2 texels (each RN(k/255)) rounded once to float;
//   * MUFU.EX2 behind __expf: restated with exp2f.
// -> pinned by the same fixture within a tolerance (tests/test_cpu.py: fraction of pixels within 1e-3 px).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

namespace {

struct F3 { float x, y, z; };
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline float unorm(uint8_t k) { return (float)k / 255.f; }   // cudaReadModeNormalizedFloat returns RN(k/255) (probed on B200)
inline float expf_dev(float x) {                               // __expf as nvcc lowers it: ex2(x*log2e) with the < -126 half/square fix-up
    float t = x * 1.4426950216293334961f;
    if (t < -126.0f) { const float r = exp2f(t * 0.5f); return r * r; }
    return exp2f(t);
}

// B-spline weights (bicubicTexture_kernel.cuh:33-57) in the contracted form of the reference build
inline void weights(float a, float* w) {
    const float sixth = 1.0f / 6.0f, a2 = a * a;
    w[0] = fmaf(a, fmaf(a, 3.0f - a, -3.0f), 1.0f) * sixth;
    w[1] = fmaf(a2, fmaf(a, 3.0f, -6.0f), 4.0f) * sixth;
    w[2] = fmaf(a, fmaf(a, fmaf(a, -3.0f, 3.0f), 3.0f), 1.0f) * sixth;
    w[3] = (a * a2) * sixth;
}
// cubicFilter (:75-84): fma(w3,c3, fma(w2,c2, fma(w0,c0, w1*c1)))
inline float cubic(const float* w, float c0, float c1, float c2, float c3) { return fmaf(w[3], c3, fmaf(w[2], c2, fmaf(w[0], c0, w[1] * c1))); }

struct Image {
    const uint8_t* p;   // [h][w][4]
    int w, h;
    F3 texel(int x, int y) const {   // clamp addressing
        const uint8_t* q = p + ((size_t)clampi(y, 0, h - 1) * w + clampi(x, 0, w - 1)) * 4;
        return F3{unorm(q[0]), unorm(q[1]), unorm(q[2])};
    }
    // point-sampled fetch at a float coordinate: floor (probed), clamp
    F3 point(float x, float y) const { return texel((int)floorf(x), (int)floorf(y)); }
    // linear-filtered fetch, unnormalised coordinates: texels around (x - 0.5, y - 0.5).  Only fractions 0 and 0.5 occur here.
    F3 linear(float x, float y) const {
        const float xs = x - 0.5f, ys = y - 0.5f;
        const float fx0 = floorf(xs), fy0 = floorf(ys);
        const double ax = xs - fx0, ay = ys - fy0;
        const int ix = (int)fx0, iy = (int)fy0;
        const F3 t00 = texel(ix, iy), t10 = texel(ix + 1, iy), t01 = texel(ix, iy + 1), t11 = texel(ix + 1, iy + 1);
        auto mix = [&](float a, float b, float c, float d) {
            return (float)((1 - ay) * ((1 - ax) * a + ax * b) + ay * ((1 - ax) * c + ax * d));
        };
        return F3{mix(t00.x, t10.x, t01.x, t11.x), mix(t00.y, t10.y, t01.y, t11.y), mix(t00.z, t10.z, t01.z, t11.z)};
    }
};

// tex2DBicubic (:88-104): x -= 0.5; px = floor(x); fx = x - px; rows py-1..py+2 filtered along x, then along y
template <bool LINEAR>
F3 bicubic(const Image& im, float x, float y) {
    x -= 0.5f; y -= 0.5f;
    const float px = floorf(x), py = floorf(y);
    float wx[4], wy[4];
    weights(x - px, wx);
    weights(y - py, wy);
    F3 row[4];
    for (int l = 0; l < 4; l++) {
        F3 t[4];
        for (int k = 0; k < 4; k++) t[k] = LINEAR ? im.linear(px - 1 + k, py - 1 + l) : im.point(px - 1 + k, py - 1 + l);
        row[l] = F3{cubic(wx, t[0].x, t[1].x, t[2].x, t[3].x), cubic(wx, t[0].y, t[1].y, t[2].y, t[3].y), cubic(wx, t[0].z, t[1].z, t[2].z, t[3].z)};
    }
    return F3{cubic(wy, row[0].x, row[1].x, row[2].x, row[3].x), cubic(wy, row[0].y, row[1].y, row[2].y, row[3].y),
              cubic(wy, row[0].z, row[1].z, row[2].z, row[3].z)};
}

inline float lum(const F3& c) { return fmaf(c.z, 0.1f, fmaf(c.x, 0.3f, c.y * 0.6f)); }   // _d_is_larger as contracted
inline float max3abs(const F3& a, const F3& b) { return fmaxf(fmaxf(fabsf(a.x - b.x), fabsf(a.y - b.y)), fabsf(a.z - b.z)); }

struct SubpixCtx {
    Image i1, i2;
    const uint8_t *c1, *c2;   // census of the x2 upsampled images, [2h][2w]
    float g[10], census[9];
    uint8_t cen(const uint8_t* c, float x, float y) const {   // point fetch, clamp, on the [2h][2w] plane
        return c[(size_t)clampi((int)floorf(y), 0, 2 * i1.h - 1) * (2 * i1.w) + clampi((int)floorf(x), 0, 2 * i1.w - 1)];
    }
};

// _d_calc_subpix_cost (:439-471) + _d_subpix_bilateral_dist (:402-421)
float subpix_cost(const SubpixCtx& s, float x1, float y1, float x2, float y2) {
    const F3 c1 = s.i1.linear(x1, y1), c2 = s.i2.linear(x2, y2);
    float cs = 0.f, ws = 0.f;
    for (int i = -9; i <= 9; i += 2)
        for (int j = -9; j <= 9; j += 2) {
            const float ii = (float)i * 0.5f, jj = (float)j * 0.5f;
            const F3 a = bicubic<true>(s.i1, x1 + jj, y1 + ii), b = bicubic<true>(s.i2, x2 + jj, y2 + ii);
            const uint8_t s1 = s.cen(s.c1, (x1 + jj) * 2.f, (y1 + ii) * 2.f), s2 = s.cen(s.c2, (x2 + jj) * 2.f, (y2 + ii) * 2.f);
            const float mod = max3abs(b, a);
            const float ad = 1.0f - expf_dev((mod * mod) / -0.010000000707805156708f);
            const float cen = s.census[__builtin_popcount((unsigned)(s1 ^ s2))];
            const float d1 = max3abs(c1, a), d2 = max3abs(c2, b);
            const float coef_r = expf_dev(fmaf(d1, d1, d2 * d2) / -0.040000002831220626831f);
            const float w = coef_r * (s.g[j < 0 ? -j : j] * s.g[i < 0 ? -i : i]);
            cs = fmaf(w, ad + cen, cs);
            ws = w + ws;
        }
    return cs / ws;
}

}  // namespace

extern "C" {

// rgba: [h][w][4] u8; out: [h_up][w_up] u8.  d_census_transform3x3_bicubic (:115-160): nine point-sampled bicubic look-ups per pixel.
void golden_census_bicubic(const uint8_t* rgba, int w, int h, int w_up, int h_up, uint8_t* out) {
    const Image im{rgba, w, h};
    const float up = (float)w / (float)w_up;   // :173
    static const int nb[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
    for (int Y = 0; Y < h_up; Y++)
        for (int X = 0; X < w_up; X++) {
            const float c = lum(bicubic<false>(im, (float)X * up, (float)Y * up));
            unsigned r = 0;
            for (int k = 0; k < 8; k++)
                if (lum(bicubic<false>(im, (float)(X + nb[k][0]) * up, (float)(Y + nb[k][1]) * up)) > c) r |= 1u << k;
            out[(size_t)Y * w_up + X] = (uint8_t)r;
        }
}

// rgba1/2: [h][w][4]; cen1_up/cen2_up: [2h][2w]; nnf: [h][w] (x,y) int16; flow: [h][w][2] in/out.  Only rows [y0, y1) are processed
// (the evaluation is 80 000 filtered fetches per pixel: tests run a band).  d_subpixel_refine (:538-634).
void golden_subpix_refine(const uint8_t* rgba1, const uint8_t* rgba2, const uint8_t* cen1_up, const uint8_t* cen2_up, const int16_t* nnf, float* flow,
                          int w, int h, int y0, int y1) {
    SubpixCtx s;
    s.i1 = Image{rgba1, w, h}; s.i2 = Image{rgba2, w, h};
    s.c1 = cen1_up; s.c2 = cen2_up;
    volatile float sig = 9.0f, lc = 0.3f;
    for (int i = 0; i < 10; i++) s.g[i] = expf(-float(i * i) / (sig * sig));                       // :701-705
    for (int i = 0; i < 9; i++) s.census[i] = 1 - expf(-float(i * i) / (lc * 8 * lc * 8));         // :707-711
    float ata[6][6];
    for (int p = 0; p < 6; p++)
        for (int q = 0; q < 6; q++) {
            float v = 0;
            for (int e = 0; e < 25; e++) {
                const float x = (float)(e % 5 - 2), y = (float)(e / 5 - 2);
                const float col[6] = {x * x, y * y, x * y, x, y, 1.f};
                v += col[p] * col[q];
            }
            ata[p][q] = v;
        }
    for (int y = y0; y < y1; y++)
        for (int x = 0; x < w; x++) {
            const int16_t Dx = nnf[((size_t)y * w + x) * 2], Dy = nnf[((size_t)y * w + x) * 2 + 1];
            if (Dx < 0 || Dy < 0 || Dx >= w || Dy >= h) continue;
            float vb[25];
            bool any = false;
            for (int dy = -2; dy <= 2; dy++)
                for (int dx = -2; dx <= 2; dx++) {
                    const float nx = (float)Dx + (float)dx / 2.f, ny = (float)Dy + (float)dy / 2.f;
                    float& o = vb[(dy + 2) * 5 + dx + 2];
                    if (nx < 0 || nx >= w || ny < 0 || ny >= h) o = 2.f;
                    else { o = subpix_cost(s, (float)x, (float)y, nx, ny); any = true; }
                }
            if (!any) continue;
            float atb[6] = {0, 0, 0, 0, 0, 0};
            for (int e = 0; e < 25; e++) {
                const float fx = (float)(e % 5 - 2), fy = (float)(e / 5 - 2);
                const float col[6] = {fx * fx, fy * fy, fx * fy, fx, fy, 1.f};
                for (int k = 0; k < 6; k++) atb[k] = fmaf(col[k], vb[e], atb[k]);
            }
            // _d_conjugate_gradient_solver (:473-536)
            float X[6], r[6], d[6], ad[6], nb2 = 0.f;
            for (int i = 0; i < 6; i++) nb2 = fmaf(atb[i], atb[i], nb2);
            const float normb = sqrtf(nb2);
            for (int i = 0; i < 6; i++) { X[i] = 0.f; r[i] = d[i] = atb[i]; }
            float rtr = normb * normb;
            int it = 0;
            while ((double)(sqrtf(rtr) / normb) > 1.0e-6 && it < 5) {
                it++;
                for (int i = 0; i < 6; i++) {
                    float acc = 0.f;
                    for (int j = 0; j < 6; j++) acc = fmaf(ata[i][j], d[j], acc);
                    ad[i] = acc;
                }
                float dad = 0.f;
                for (int i = 0; i < 6; i++) dad = fmaf(d[i], ad[i], dad);
                const float alpha = rtr / dad;
                for (int i = 0; i < 6; i++) { X[i] = fmaf(alpha, d[i], X[i]); r[i] = fmaf(ad[i], -alpha, r[i]); }
                const float rtrold = rtr;
                rtr = 0.f;
                for (int i = 0; i < 6; i++) rtr = fmaf(r[i], r[i], rtr);
                const float beta = rtr / rtrold;
                for (int i = 0; i < 6; i++) d[i] = fmaf(beta, d[i], r[i]);
            }
            const float den = fmaf(X[2], X[2], (X[0] * -4.0f) * X[1]);   // :620
            if (den == 0.f) continue;
            const float subx = fmaf(X[3] + X[3], X[1], -(X[4] * X[2])) / den, suby = fmaf(X[0] + X[0], X[4], -(X[3] * X[2])) / den;
            if (fabsf(suby) <= 3.f && fabsf(subx) <= 3.f) {
                flow[((size_t)y * w + x) * 2] = fmaf((float)(Dx - x), 2.0f, subx) * 0.5f;
                flow[((size_t)y * w + x) * 2 + 1] = fmaf((float)(Dy - y), 2.0f, suby) * 0.5f;
            }
        }
}

}  // extern "C"
