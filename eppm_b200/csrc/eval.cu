// Evaluation of a flow field against ground truth on the device, and the .flo wire format.
//   eppm_eval_flow : bao_calc_flow_error (basic/bao_flow_tools.cpp:64-111: mean end-point error and mean angular error over the pixels whose
//                    ground truth is known and non-zero, inside an optional border) and bao_calc_flow_error_percentage (:114-141: share of
//                    the pixels with known ground truth whose end-point error exceeds a threshold), one launch pair for a whole batch.
//   eppm_write_flo / eppm_read_flo : Middlebury .flo ("PIEH", int32 width, int32 height, rows of interleaved (u, v) float32, little endian;
//                    3rdparty/middlebury/flowIO.cpp:122-160, 56-116) straight from / to the library's interleaved flow layout.
// Per-pixel arithmetic is the reference's (float sqrt / acos); the SUMS are accumulated in double in a fixed order (per-block partials,
// then one thread block folds them), so the result is deterministic; the reference adds floats sequentially on one host thread, which a
// parallel reduction cannot reproduce bit for bit -- the tests compare within the rounding error of that float sum.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "eppm_internal.h"

namespace eppm {

constexpr int EV_BLOCKS = 256;   // partial sums per pair

struct EvalPartial {
    double epe, aae;
    unsigned long long n_err, n_known, n_bad;
};

__global__ void __launch_bounds__(256) k_eval_partial(const float2* __restrict__ flow, const float2* __restrict__ gt, int w, int h, int border, float thresh,
                                                      EvalPartial* __restrict__ part) {
    const size_t off = (size_t)blockIdx.y * w * h;
    double epe = 0.0, aae = 0.0;
    unsigned long long n_err = 0, n_known = 0, n_bad = 0;
    const int total = w * h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / w, x = i - y * w;
        const float2 g = gt[off + i], f = flow[off + i];
        const float agu = fabsf(g.x), agv = fabsf(g.y);
        const float du = __fsub_rn(f.x, g.x), dv = __fsub_rn(f.y, g.y);
        const float e = sqrtf(__fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv)));   // :84
        if (agu <= EPPM_UNKNOWN_FLOW_THRESH || agv <= EPPM_UNKNOWN_FLOW_THRESH) {   // :126 known ground truth
            n_known++;
            if (!(e <= thresh)) n_bad++;                                           // :132
        }
        if (y >= border && y < h - border && x >= border && x < w - border &&
            ((agu > 0.f && agu <= EPPM_UNKNOWN_FLOW_THRESH) || (agv > 0.f && agv <= EPPM_UNKNOWN_FLOW_THRESH))) {   // :76
            n_err++;
            const float num = __fadd_rn(__fadd_rn(__fmul_rn(f.x, g.x), __fmul_rn(f.y, g.y)), 1.0f);
            const float den = __fmul_rn(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(f.x, f.x), __fmul_rn(f.y, f.y)), 1.0f)),
                                        sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(g.x, g.x), __fmul_rn(g.y, g.y)), 1.0f)));
            aae += (double)acosf(__fdiv_rn(num, den));   // :81-82
            epe += (double)e;
        }
    }
    __shared__ double s_e[256], s_a[256];
    __shared__ unsigned long long s_n[3][256];
    s_e[threadIdx.x] = epe; s_a[threadIdx.x] = aae;
    s_n[0][threadIdx.x] = n_err; s_n[1][threadIdx.x] = n_known; s_n[2][threadIdx.x] = n_bad;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {   // fixed tree: deterministic
        if (threadIdx.x < s) {
            s_e[threadIdx.x] += s_e[threadIdx.x + s]; s_a[threadIdx.x] += s_a[threadIdx.x + s];
            for (int k = 0; k < 3; k++) s_n[k][threadIdx.x] += s_n[k][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = EvalPartial{s_e[0], s_a[0], s_n[0][0], s_n[1][0], s_n[2][0]};
}

__global__ void k_eval_final(const EvalPartial* __restrict__ part, int n_part, eppm_flow_error* __restrict__ out) {
    if (threadIdx.x != 0) return;
    const EvalPartial* p = part + (size_t)blockIdx.x * n_part;
    double epe = 0.0, aae = 0.0;
    unsigned long long n_err = 0, n_known = 0, n_bad = 0;
    for (int i = 0; i < n_part; i++) { epe += p[i].epe; aae += p[i].aae; n_err += p[i].n_err; n_known += p[i].n_known; n_bad += p[i].n_bad; }
    eppm_flow_error r;
    r.epe = n_err ? epe / (double)n_err : 0.0;
    r.aae_deg = n_err ? aae / (double)n_err * 180.0 / (double)3.14159f : 0.0;   // :96 uses the literal 3.14159f
    r.outlier_frac = n_known ? 1.0 - (double)(n_known - n_bad) / (double)n_known : 0.0;
    r.n_valid = (long long)n_err;
    r.n_known = (long long)n_known;
    out[blockIdx.x] = r;
}

}  // namespace eppm

using namespace eppm;

extern "C" {

int eppm_eval_flow(eppm_context* c, const float* d_flow, const float* d_gt, int n, int border, float outlier_thresh, eppm_flow_error* out) {
    if (!c || !d_flow || !d_gt || !out || n < 1 || border < 0) { set_error("eppm_eval_flow: bad argument"); return EPPM_ERR_ARG; }
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != c->device) cudaSetDevice(c->device);
    EvalPartial* part = nullptr;
    eppm_flow_error* d_out = nullptr;
    int rc = EPPM_OK;
    if (!cuda_ok(cudaMalloc((void**)&part, (size_t)n * EV_BLOCKS * sizeof(EvalPartial)), "cudaMalloc") ||
        !cuda_ok(cudaMalloc((void**)&d_out, (size_t)n * sizeof(eppm_flow_error)), "cudaMalloc")) {
        rc = EPPM_ERR_CUDA;
    } else {
        k_eval_partial<<<dim3(EV_BLOCKS, n), 256, 0, c->stream>>>(reinterpret_cast<const float2*>(d_flow), reinterpret_cast<const float2*>(d_gt), c->w, c->h, border,
                                                                  outlier_thresh, part);
        k_eval_final<<<n, 32, 0, c->stream>>>(part, EV_BLOCKS, d_out);
        EPPM_LAUNCH_COUNT(2);
        if (!cuda_ok(cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(eppm_flow_error), cudaMemcpyDeviceToHost, c->stream), "eval copy") ||
            !cuda_ok(cudaStreamSynchronize(c->stream), "eppm_eval_flow"))
            rc = EPPM_ERR_CUDA;
    }
    if (part) cudaFree(part);
    if (d_out) cudaFree(d_out);
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
    return rc;
}

int eppm_write_flo(const char* path, const float* flow_uv, int h, int w) {
    if (!path || !flow_uv || h < 1 || w < 1) { set_error("eppm_write_flo: bad argument"); return EPPM_ERR_ARG; }
    const char* dot = strrchr(path, '.');
    if (!dot || strcmp(dot, ".flo") != 0) { set_error("eppm_write_flo: the file name needs the extension .flo (flowIO.cpp:131-135)"); return EPPM_ERR_ARG; }
    FILE* f = fopen(path, "wb");
    if (!f) { set_error(std::string("eppm_write_flo: cannot open ") + path); return EPPM_ERR_ARG; }
    const int32_t hdr[2] = {w, h};
    bool ok = fwrite("PIEH", 1, 4, f) == 4 && fwrite(hdr, 4, 2, f) == 2 && fwrite(flow_uv, sizeof(float) * 2, (size_t)w * h, f) == (size_t)w * h;
    ok = (fclose(f) == 0) && ok;
    if (!ok) set_error(std::string("eppm_write_flo: short write to ") + path);
    return ok ? EPPM_OK : EPPM_ERR_STATE;
}

int eppm_read_flo(const char* path, float* flow_uv, int* h, int* w, size_t capacity_floats) {
    if (!path || !h || !w) { set_error("eppm_read_flo: bad argument"); return EPPM_ERR_ARG; }
    FILE* f = fopen(path, "rb");
    if (!f) { set_error(std::string("eppm_read_flo: cannot open ") + path); return EPPM_ERR_ARG; }
    char tag[4];
    int32_t hdr[2];
    int rc = EPPM_OK;
    if (fread(tag, 1, 4, f) != 4 || memcmp(tag, "PIEH", 4) != 0 || fread(hdr, 4, 2, f) != 2 || hdr[0] < 1 || hdr[1] < 1 || hdr[0] > 99999 || hdr[1] > 99999) {
        set_error("eppm_read_flo: not a .flo file (tag PIEH, sane width / height: flowIO.cpp:70-90)");
        rc = EPPM_ERR_ARG;
    } else {
        *w = hdr[0]; *h = hdr[1];
        const size_t nfl = (size_t)hdr[0] * hdr[1] * 2;
        if (flow_uv) {   // a null buffer only queries the size
            if (capacity_floats < nfl) { set_error("eppm_read_flo: buffer too small"); rc = EPPM_ERR_ARG; }
            else if (fread(flow_uv, sizeof(float), nfl, f) != nfl) { set_error("eppm_read_flo: file is too short"); rc = EPPM_ERR_STATE; }
            else if (fgetc(f) != EOF) { set_error("eppm_read_flo: file is too long"); rc = EPPM_ERR_STATE; }
        }
    }
    fclose(f);
    return rc;
}

}  // extern "C"
