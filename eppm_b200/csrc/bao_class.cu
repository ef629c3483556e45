// bao_flow_patchmatch_multiscale_cuda — the reference's public host class (bao_flow_patchmatch_multiscale_cuda.h:33-45),
// implemented over the batched eppm_* context with a batch of one pair.  Same call sequence and memory convention:
//   init(h,w) once per resolution -> per pair set_data(img1,img2) -> compute_flow(u,v[,color]).
// set_data uploads the pair and runs "prepare" (…cuda.cpp:159-168 + _prepare_data :212-215); compute_flow runs the rest
// and fills u[y][x], v[y][x] (…cuda.cpp:217-306).  Errors are reported like the reference does: a message on stderr.
#include <stdio.h>
#include <math.h>

#include "../../include/compat/bao_flow_patchmatch_multiscale_cuda.h"
#include "eppm_internal.h"

static void report(const char* where) { fprintf(stderr, "EPPM(b200) error in %s: %s\n", where, eppm_last_error()); }

bao_flow_patchmatch_multiscale_cuda::bao_flow_patchmatch_multiscale_cuda() : m_ctx(NULL), m_h(0), m_w(0), m_d_flow(NULL), m_h_flow(NULL), m_has_data(false) {
    m_d_rgb[0] = m_d_rgb[1] = NULL;
}

bao_flow_patchmatch_multiscale_cuda::~bao_flow_patchmatch_multiscale_cuda() {
    if (m_h_flow) cudaFreeHost(m_h_flow);
    if (m_ctx) eppm_destroy(m_ctx);
}

void bao_flow_patchmatch_multiscale_cuda::init(int h, int w) {
    if (m_ctx) {  // the reference leaks on re-init (…cuda.cpp:112-157); here the old context is released
        eppm_destroy(m_ctx);
        m_ctx = NULL;
    }
    if (m_h_flow) { cudaFreeHost(m_h_flow); m_h_flow = NULL; }
    m_h = h; m_w = w; m_has_data = false;
    int dev = 0;
    cudaGetDevice(&dev);
    if (eppm_create(&m_ctx, dev, h, w, 1, NULL) != EPPM_OK) { report("init"); m_ctx = NULL; return; }
    m_d_rgb[0] = m_ctx->d_rgb[0];
    m_d_rgb[1] = m_ctx->d_rgb[1];
    m_d_flow = m_ctx->d_flow_out;
    cudaMallocHost((void**)&m_h_flow, (size_t)h * w * 2 * sizeof(float));
}

void bao_flow_patchmatch_multiscale_cuda::init(unsigned char*** img1, unsigned char*** img2, int h, int w) {
    init(h, w);
    set_data(img1, img2);
}

bool bao_flow_patchmatch_multiscale_cuda::set_data(unsigned char*** img1, unsigned char*** img2) {
    if (!m_ctx) { fprintf(stderr, "EPPM(b200): set_data before init\n"); return true; }
    const size_t bytes = (size_t)m_h * m_w * 3;
    cudaStream_t s = (cudaStream_t)eppm_stream(m_ctx);
    // img[0][0] is the contiguous h*w*3 block of the bao_alloc layout (basic/bao_basic.h:144-162)
    cudaMemcpyAsync(m_d_rgb[0], img1[0][0], bytes, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(m_d_rgb[1], img2[0][0], bytes, cudaMemcpyHostToDevice, s);
    if (eppm_stage_prepare(m_ctx, m_d_rgb[0], m_d_rgb[1], 1) != EPPM_OK) report("set_data");
    // host buffers may be reused by the caller as soon as we return
    if (eppm_synchronize(m_ctx) != EPPM_OK) report("set_data");
    m_has_data = true;
    return true;
}

// Middlebury colour wheel (the reference's optional color_flow output, basic/bao_basic_cuda.cuh:745-849 /
// 3rdparty/middlebury/colorcode.cpp): computed on the host from the returned flow; not on the hot path.
static void flow_to_color(const float* uv, unsigned char*** out, int h, int w, float max_x, float max_y) {
    static int wheel[55][3];
    static int ncols = 0;
    if (!ncols) {
        const int RY = 15, YG = 6, GC = 4, CB = 11, BM = 13, MR = 6;
        int k = 0;
        for (int i = 0; i < RY; i++, k++) { wheel[k][0] = 255; wheel[k][1] = 255 * i / RY; wheel[k][2] = 0; }
        for (int i = 0; i < YG; i++, k++) { wheel[k][0] = 255 - 255 * i / YG; wheel[k][1] = 255; wheel[k][2] = 0; }
        for (int i = 0; i < GC; i++, k++) { wheel[k][0] = 0; wheel[k][1] = 255; wheel[k][2] = 255 * i / GC; }
        for (int i = 0; i < CB; i++, k++) { wheel[k][0] = 0; wheel[k][1] = 255 - 255 * i / CB; wheel[k][2] = 255; }
        for (int i = 0; i < BM; i++, k++) { wheel[k][0] = 255 * i / BM; wheel[k][1] = 0; wheel[k][2] = 255; }
        for (int i = 0; i < MR; i++, k++) { wheel[k][0] = 255; wheel[k][1] = 0; wheel[k][2] = 255 - 255 * i / MR; }
        ncols = k;
    }
    const float maxrad = sqrtf(max_x * max_x + max_y * max_y);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float ux = uv[2 * ((size_t)y * w + x)], uy = uv[2 * ((size_t)y * w + x) + 1];
            if (!(fabsf(ux) < 999999.f && fabsf(uy) < 999999.f)) {   // unknown flow stays black (basic/bao_basic_cuda.cuh:823)
                out[y][x][0] = out[y][x][1] = out[y][x][2] = 0;
                continue;
            }
            const float fx = ux / maxrad, fy = uy / maxrad;
            const float rad = sqrtf(fx * fx + fy * fy);
            const float a = atan2f(-fy, -fx) / 3.14159f;   // the reference's constant (:781)
            const float fk = (a + 1.0f) / 2.0f * (ncols - 1);
            const int k0 = (int)fk, k1 = (k0 + 1) % ncols;
            const float f = fk - k0;
            for (int b = 0; b < 3; b++) {
                float col = (1 - f) * wheel[k0][b] / 255.0f + f * wheel[k1][b] / 255.0f;
                col = rad <= 1 ? 1 - rad * (1 - col) : col * 0.75f;
                out[y][x][b] = (unsigned char)(255.0f * col);
            }
        }
}

void bao_flow_patchmatch_multiscale_cuda::compute_flow(float** disp1_x, float** disp1_y, unsigned char*** color_flow) {
    if (!m_ctx || !m_has_data) { fprintf(stderr, "EPPM(b200): compute_flow before init/set_data\n"); return; }
    cudaStream_t s = (cudaStream_t)eppm_stream(m_ctx);
    if (eppm_stage_patchmatch(m_ctx) != EPPM_OK || eppm_stage_consistency(m_ctx) != EPPM_OK || eppm_stage_c2f(m_ctx, m_d_flow) != EPPM_OK) {
        report("compute_flow");
        return;
    }
    const size_t n = (size_t)m_h * m_w;
    cudaMemcpyAsync(m_h_flow, m_d_flow, n * 2 * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (eppm_synchronize(m_ctx) != EPPM_OK) { report("compute_flow"); return; }
    float* u = disp1_x[0];  // contiguous h*w block (basic/bao_basic.h:124-133)
    float* v = disp1_y[0];
    for (size_t i = 0; i < n; i++) { u[i] = m_h_flow[2 * i]; v[i] = m_h_flow[2 * i + 1]; }
    if (color_flow) flow_to_color(m_h_flow, color_flow, m_h, m_w, 20.f, 20.f);  // …cuda.cpp:311
}
