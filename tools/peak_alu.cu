// ALU peaks of the box's B200 (SURVEY.md §8d asks for them: MEASURED_PEAKS.json holds only HBM and bf16 tensor figures).
// Development tool, run under gpurun:   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/peak_alu tools/peak_alu.cu && build/peak_alu
// Writes gpurun_out/alu_peaks.json: warp-instructions per clock per SM and lane-ops/s for
//   ffma      scalar FP32 FMA            (the nominal 128 lanes/SM/clk = 4 warp-instr/clk/SM)
//   ffma2     packed FP32x2 FMA          (fma.rn.f32x2: does it double the lane rate or halve the issue rate?)
//   mufu      MUFU.EX2                   (nominal 16 lanes/SM/clk)
//   popc      POPC                       (shares the XU pipe with MUFU on this architecture?)
//   fmnmx3    3-input FMNMX              (ALU pipe)
//   mix       the instruction mix of one patch sample of the refine kernel (packed form): what the issue port sustains on it
// Every kernel keeps 8 independent dependency chains per thread so that latency never limits; 8 CTAs x 256 threads per SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2(float t) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t)); return r; }

constexpr int ITERS = 4096, CH = 8;

__global__ void k_ffma(float* out, float a, float b) {
    float v[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) v[k] = threadIdx.x * 1e-3f + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < CH; k++) v[k] = __fmaf_rn(v[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CH; k++) s += v[k];
    if (s == 123.456f) out[0] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
    f32x2 v[CH];
    const f32x2 A = pk2(a, a), B = pk2(b, b);
#pragma unroll
    for (int k = 0; k < CH; k++) v[k] = pk2(threadIdx.x * 1e-3f + k, 1.f + k);
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < CH; k++) v[k] = fma2(v[k], A, B);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CH; k++) { float x, y; upk2(v[k], x, y); s += x + y; }
    if (s == 123.456f) out[0] = s;
}
__global__ void k_mufu(float* out, float a) {
    float v[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) v[k] = -(threadIdx.x * 1e-3f + k) * a;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < CH; k++) v[k] = ex2(v[k]);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CH; k++) s += v[k];
    if (s == 123.456f) out[0] = s;
}
__global__ void k_popc(float* out, unsigned a) {
    unsigned v[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) v[k] = threadIdx.x * 2654435761u + k * a;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < CH; k++) v[k] = __popc(v[k]) + a;   // POPC + IADD: the add is on the ALU pipe; halve its share below
    }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) s += v[k];
    if (s == 0xdeadbeefu) out[0] = (float)s;
}
__global__ void k_iadd(float* out, unsigned a) {   // baseline for k_popc: the same loop with the POPC replaced by a second add
    unsigned v[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) v[k] = threadIdx.x * 2654435761u + k * a;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < CH; k++) v[k] = (v[k] ^ a) + a;
    }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) s += v[k];
    if (s == 0xdeadbeefu) out[0] = (float)s;
}
__global__ void k_fmnmx3(float* out, float a, float b) {
    float v[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) v[k] = threadIdx.x * 1e-3f + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < CH; k++) v[k] = fmaxf(fmaxf(fabsf(v[k]), fabsf(a)), fabsf(b));
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CH; k++) s += v[k];
    if (s == 123.456f) out[0] = s;
}
// one "sample pair" of the packed refine kernel without its loads: 4 FADD2 + 4 FADD + 4 FMNMX3 + 3 packed squares + 8 packed
// division/log2e ops + 4 MUFU + 2 LOP3 + 2 POPC + 2 packed (1 - e, + census) + 2 FMUL + 2 packed accumulations = 37 issue slots
__global__ void k_mix(float* out, float a, float b, unsigned m) {
    f32x2 cs[CH / 2], ws[CH / 2];
    float x[CH / 2], y[CH / 2];
#pragma unroll
    for (int k = 0; k < CH / 2; k++) { cs[k] = ws[k] = pk2(0.f, 0.f); x[k] = threadIdx.x * 1e-3f + k; y[k] = x[k] * a; }
    const f32x2 R = pk2(-99.99999237060546875f, -99.99999237060546875f), D = pk2(0.010000000707805156708f, 0.010000000707805156708f), Z = pk2(0.f, 0.f);
    const f32x2 L2E = pk2(1.4426950216293334961f, 1.4426950216293334961f), ONE = pk2(1.f, 1.f);
    for (int i = 0; i < ITERS / 8; i++) {
#pragma unroll
        for (int k = 0; k < CH / 2; k++) {
            const f32x2 p1 = pk2(x[k], y[k]), p2 = pk2(y[k], a), c2 = pk2(b, x[k]);
            float d0, d1, e0, e1, f0, f1, g0, g1;
            upk2(add2(p1, p2), d0, d1); upk2(add2(p1, c2), e0, e1); upk2(add2(c2, p2), f0, f1); upk2(add2(c2, p1), g0, g1);
            const float ca = fmaxf(fmaxf(fabsf(d0), fabsf(d1)), fabsf(x[k] - a)), cb = fmaxf(fmaxf(fabsf(e0), fabsf(e1)), fabsf(y[k] - a));
            const float da = fmaxf(fmaxf(fabsf(f0), fabsf(f1)), fabsf(x[k] - b)), db = fmaxf(fmaxf(fabsf(g0), fabsf(g1)), fabsf(y[k] - b));
            f32x2 c = pk2(ca, cb), d2 = pk2(da, db);
            f32x2 xc = fma2(c, c, Z), xw = fma2(ONE, ONE, fma2(d2, d2, Z));
            f32x2 qc0 = fma2(xc, R, Z), qw0 = fma2(xw, R, Z);
            f32x2 qc = fma2(R, fma2(qc0, D, xc), qc0), qw = fma2(R, fma2(qw0, D, xw), qw0);
            float t1a, t1b, t2a, t2b;
            upk2(fma2(qc, L2E, Z), t1a, t1b); upk2(fma2(qw, L2E, Z), t2a, t2b);
            const f32x2 e = pk2(ex2(t1a), ex2(t1b));
            const unsigned ua = __popc(__float_as_uint(ca) ^ m), ub = __popc(__float_as_uint(cb) ^ m);
            const f32x2 ct = add2(add2(ONE, e), pk2(__uint_as_float(ua), __uint_as_float(ub)));
            const f32x2 w = pk2(ex2(t2a) * a, ex2(t2b) * a);
            cs[k] = fma2(ct, w, cs[k]);
            ws[k] = add2(ws[k], w);
            upk2(w, x[k], y[k]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CH / 2; k++) { float p, q; upk2(cs[k], p, q); s += p + q; upk2(ws[k], p, q); s += p + q; }
    if (s == 123.456f) out[0] = s;
}

template <class F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, ctas = sms * 8, thr = 256;
    float* d; cudaMalloc(&d, 64);
    const double warps = (double)ctas * thr / 32;
    const double n_simple = warps * ITERS * CH;   // warp-instructions of the measured kind per launch
    struct { const char* name; double ms; double winstr; double lanes_per_instr; } r[7];
    r[0] = {"ffma", time_ms([&] { k_ffma<<<ctas, thr>>>(d, 1.0001f, 0.5f); }), n_simple, 32};
    r[1] = {"ffma2", time_ms([&] { k_ffma2<<<ctas, thr>>>(d, 1.0001f, 0.5f); }), n_simple, 64};
    r[2] = {"mufu_ex2", time_ms([&] { k_mufu<<<ctas, thr>>>(d, 0.001f); }), n_simple, 32};
    r[3] = {"popc_plus_iadd", time_ms([&] { k_popc<<<ctas, thr>>>(d, 3u); }), n_simple, 32};
    r[4] = {"lop3_plus_iadd", time_ms([&] { k_iadd<<<ctas, thr>>>(d, 3u); }), n_simple, 32};
    r[5] = {"fmnmx3", time_ms([&] { k_fmnmx3<<<ctas, thr>>>(d, 0.25f, 0.125f); }), n_simple, 32};
    r[6] = {"refine_sample_mix_37slots", time_ms([&] { k_mix<<<ctas, thr>>>(d, 0.9f, 0.3f, 0x0f0f0f0fu); }), warps * (ITERS / 8) * (CH / 2) * 37.0, 32};
    cudaError_t e = cudaDeviceSynchronize();
    system("mkdir -p gpurun_out");
    FILE* f = fopen("gpurun_out/alu_peaks.json", "w");
    if (!f) { perror("gpurun_out/alu_peaks.json"); return 1; }
    fprintf(f, "{\n \"gpu\": \"%s\", \"sms\": %d, \"attr_clock_mhz\": %.0f, \"cuda_error\": %d,\n", p.name, sms, clk_khz / 1e3, (int)e);
    fprintf(f, " \"method\": \"8 CTAs x 256 threads per SM, 8 independent chains per thread, best of 5 launches, CUDA events; per_clk figures use attr_clock_mhz\",\n");
    for (int i = 0; i < 7; i++) {
        const double per_s = r[i].winstr / (r[i].ms * 1e-3);
        fprintf(f, " \"%s\": {\"ms\": %.4f, \"warp_instr_per_s\": %.4e, \"warp_instr_per_clk_per_sm\": %.3f, \"lane_ops_per_s\": %.4e}%s\n", r[i].name, r[i].ms, per_s,
                per_s / (clk_khz * 1e3) / sms, per_s * r[i].lanes_per_instr, i < 6 ? "," : "");
        printf("%-28s %8.3f ms  %.3f warp-instr/clk/SM  %.3e lane-ops/s\n", r[i].name, r[i].ms, per_s / (clk_khz * 1e3) / sms, per_s * r[i].lanes_per_instr);
    }
    fprintf(f, "}\n");
    fclose(f);
    return e == cudaSuccess ? 0 : 1;
}
