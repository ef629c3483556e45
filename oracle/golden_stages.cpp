// TEST INFRASTRUCTURE ONLY — same scope and rules as golden.cpp (see golden.h): only tests/ may call this.
//
// Single-threaded CPU restatement of the stage functions the reference's host class declares (bao_flow_patchmatch_multiscale_cuda.cpp:40-62)
// but compute_flow never calls; dense [h][w] planes, images as [h][w][4] u8.  Each function cites the reference lines it follows.
// Pinned by tests/golden/refstage_*.npz (outputs of the reference build on a B200, tools/gen_golden_stages.py): the integer / IEEE
// ones bit-exact, the ones that go through __expf (MUFU.EX2) within the tolerance stated in tests/test_cpu.py.
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

namespace {

const int INVALID_LOCATION = -10000;   // bao_pmflow_refine_kernel.cu:46
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline float unorm(uint8_t k) { return (float)k / 255.f; }
inline float expf_dev(float x) {       // __expf as lowered without -ftz: ex2(x*log2e) with the < -126 half/square fix-up; MUFU.EX2 ~ exp2f
    float t = x * 1.4426950216293334961f;
    if (t < -126.0f) { const float r = exp2f(t * 0.5f); return r * r; }
    return exp2f(t);
}
struct F3 { float x, y, z; };
inline F3 texel(const uint8_t* img, int w, int h, int x, int y) {   // clamp addressing, normalised-float read
    const uint8_t* q = img + ((size_t)clampi(y, 0, h - 1) * w + clampi(x, 0, w - 1)) * 4;
    return F3{unorm(q[0]), unorm(q[1]), unorm(q[2])};
}
inline float max3abs(const F3& a, const F3& b) { return fmaxf(fmaxf(fabsf(a.x - b.x), fabsf(a.y - b.y)), fabsf(a.z - b.z)); }

// `short(float)` on the device (pinned by the fixture): cvt.rzi.s32.f32 -- truncate, saturate to 32 bits, NaN -> 0 -- then the low 16 bits
inline int16_t f2s16(float v) {
    int32_t i;
    if (v != v) i = 0;
    else if (v >= 2147483648.f) i = INT32_MAX;
    else if (v <= -2147483648.f) i = INT32_MIN;
    else i = (int32_t)v;
    return (int16_t)(uint16_t)((uint32_t)i & 0xffffu);
}

// d_left_right_check_buffered (bao_pmflow_refine_kernel.cu:93-122)
void lr_buffered_pass(int16_t* out_nnf, float* out_cost, const int16_t* nnf, const float* cost, const int16_t* nnf2, int w, int h) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const size_t id = (size_t)y * w + x;
            const int16_t dx = nnf[2 * id], dy = nnf[2 * id + 1];
            int16_t ox = INVALID_LOCATION, oy = INVALID_LOCATION;
            float oc = FLT_MAX;
            if (!(dy < 0 || dy >= h || dx < 0 || dx >= w)) {
                const size_t t = (size_t)dy * w + dx;
                if (!(abs(nnf2[2 * t] - x) > 50 || abs(nnf2[2 * t + 1] - y) > 50)) { ox = dx; oy = dy; oc = cost[id]; }
            }
            out_nnf[2 * id] = ox; out_nnf[2 * id + 1] = oy; out_cost[id] = oc;
        }
}

// _d_bilateral_weight (bao_pmflow_refine_kernel.cu:756-762) with POSTPROC_BLF_SIG_R = 0.02, spatial part G[|dx|]*G[|dy|]
struct Blf {
    float g[11];
    Blf() {
        volatile int s = 5;   // POSTPROC_BLF_SIG_S (defs.h:64)
        for (int i = 0; i <= 10; i++) g[i] = expf(-float(i * i) / float(s * s));
    }
    float weight(const F3& c, const F3& p, int adx, int ady) const {
        const float dr = max3abs(p, c);
        const float coef_r = expf_dev((dr * dr) / -(0.02f * 0.02f));
        return coef_r * (g[adx] * g[ady]);
    }
};

}  // namespace

extern "C" {

// baoCudaLeftRightCheck_Buffered (:124-140): forward into temporaries, backward in place against the ORIGINAL forward field, copy back
void golden_lr_check_buffered(int16_t* nnf, float* cost, int16_t* nnf2, float* cost2, int w, int h) {
    std::vector<int16_t> tn((size_t)w * h * 2);
    std::vector<float> tc((size_t)w * h);
    lr_buffered_pass(tn.data(), tc.data(), nnf, cost, nnf2, w, h);
    std::vector<int16_t> n2((size_t)w * h * 2);
    std::vector<float> c2((size_t)w * h);
    lr_buffered_pass(n2.data(), c2.data(), nnf2, cost2, nnf, w, h);   // every pixel reads its own entry and the untouched forward field
    memcpy(nnf2, n2.data(), n2.size() * 2); memcpy(cost2, c2.data(), c2.size() * 4);
    memcpy(nnf, tn.data(), tn.size() * 2); memcpy(cost, tc.data(), tc.size() * 4);
}

// d_convert_flow_to_nnf (:657-676)
void golden_flow_to_nnf(const float* flow, int16_t* nnf, int w, int h) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const size_t id = (size_t)y * w + x;
            const float fx = flow[2 * id], fy = flow[2 * id + 1];
            if (fx > 1e9f || fy > 1e9f) { nnf[2 * id] = nnf[2 * id + 1] = (int16_t)INVALID_LOCATION; continue; }
            nnf[2 * id] = f2s16(fx + (float)x);
            nnf[2 * id + 1] = f2s16(fy + (float)y);
        }
}

// d_flow_cutoff (:891-900) with the reference's __min / __max macros
void golden_flow_cutoff(float* flow, int w, int h, float m) {
    for (size_t i = 0; i < (size_t)w * h * 2; i++) {
        const float v = flow[i];
        const float a = (m < v) ? m : v;
        flow[i] = (-m > a) ? -m : a;
    }
}

// d_eliminate_still_region_flow + _d_compute_patch_dist_ad_L2 (bao_pmflow_kernel.cu:555-586, 2071-2081)
void golden_eliminate_still(float* flow, const uint8_t* rgba1, const uint8_t* rgba2, int w, int h) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float cs = 0.f, ws = 0.f;
            for (int i = -9; i <= 9; i += 2)
                for (int j = -9; j <= 9; j += 2) {
                    const float c = max3abs(texel(rgba1, w, h, x + j, y + i), texel(rgba2, w, h, x + j, y + i));
                    cs += 1.0f - expf_dev((c * c) / -0.010000000707805156708f);
                    ws += 1.0f;
                }
            if ((double)(cs / ws) <= 0.1) flow[2 * ((size_t)y * w + x)] = flow[2 * ((size_t)y * w + x) + 1] = 0.f;
        }
}

// d_image_bilateral_filtering (bao_pmflow_refine_kernel.cu:976-1020); out = [h][w][3] (the reference leaves alpha uninitialised)
void golden_image_smoothing(const uint8_t* rgba, uint8_t* out_rgb, int w, int h) {
    static const Blf blf;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const F3 c = texel(rgba, w, h, x, y);
            float n[3] = {0, 0, 0}, ws = 0.f;
            for (int dy = -10; dy <= 10; dy++)
                for (int dx = -10; dx <= 10; dx++) {
                    const int cx = x + dx, cy = y + dy;
                    if (cx < 0 || cy < 0 || cx >= w || cy >= h) continue;
                    const float wt = blf.weight(c, texel(rgba, w, h, cx, cy), abs(dx), abs(dy));
                    const uint8_t* q = rgba + ((size_t)cy * w + cx) * 4;
                    for (int k = 0; k < 3; k++) n[k] = fmaf(wt, (float)q[k], n[k]);
                    ws += wt;
                }
            uint8_t* o = out_rgb + ((size_t)y * w + x) * 3;
            for (int k = 0; k < 3; k++) o[k] = ws != 0.f ? (uint8_t)(unsigned)(n[k] / ws) : rgba[((size_t)y * w + x) * 4 + k];
        }
}

// d_bilateral_upsample_flow (:829-865): out is left untouched where no tap contributes
void golden_flow_bilateral_upsample(float* out, const uint8_t* rgba, int w, int h, const float* small, int ws_, float ratio) {
    static const Blf blf;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const F3 c = texel(rgba, w, h, x, y);
            float nx = 0.f, ny = 0.f, wsum = 0.f;
            for (int dy = -10; dy <= 10; dy++)
                for (int dx = -10; dx <= 10; dx++) {
                    const int cx = x + dx, cy = y + dy;
                    if (cx < 0 || cy < 0 || cx >= w || cy >= h) continue;
                    const float* f = small + 2 * ((size_t)(int)((float)cy / ratio) * ws_ + (int)((float)cx / ratio));
                    if (f[0] > 1e9f || f[1] > 1e9f) continue;
                    const float wt = blf.weight(c, texel(rgba, w, h, cx, cy), abs(dx), abs(dy));
                    nx = fmaf(wt, f[0], nx); ny = fmaf(wt, f[1], ny); wsum += wt;
                }
            if (wsum != 0.f) {
                out[2 * ((size_t)y * w + x)] = (nx / wsum) * ratio;
                out[2 * ((size_t)y * w + x) + 1] = (ny / wsum) * ratio;
            }
        }
}

}  // extern "C"
