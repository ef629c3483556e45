// Hardware probe (development tool, run under gpurun): pins the device-defined arithmetic the
// product relies on and writes it to gpurun_out/probe_hw.json.
//  1. __expf Gaussian tap weights of the three pyramid blurs (basic/bao_basic_cuda.cuh:437-467 of the reference
//     evaluates __expf(-(dy^2+dx^2)/(2 sigma^2)) per tap) -> bit patterns for the CPU oracle.
//  2. constant-divisor division: q = x*r; rem = fma(q, 0.01', x); q' = fma(r, rem, q) against x / -0.01'
//     exhaustively for every float x in [2^-20, 4).
//  3. u8 -> float as k * (1/255) variants against k / 255.f.
//  4. a sample of ex2.approx outputs (for the tolerance study of the CPU oracle).
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__global__ void k_gauss(float* out, float sigma2x2, int r) {
    int n = 2 * r + 1;
    int i = threadIdx.x;
    if (i >= n * n) return;
    int dy = i / n - r, dx = i % n - r;
    out[i] = __expf(-(float)(dy * dy + dx * dx) / sigma2x2);
}

__device__ __forceinline__ float div_const_fast(float x) {
    // sequence nvcc 12.9 emits for x / -(0.1f*0.1f): r' = fma(fma(-R33, R30, 1), -R33, -R33)
    const float R33 = 99.99999237060546875f;  // 0x42c7ffff
    const float R30 = 0.010000000707805156708f; // 0x3c23d70b
    float t = __fmaf_rn(-R33, R30, 1.0f);
    float rr = __fmaf_rn(t, -R33, -R33);
    float q0 = __fmaf_rn(x, rr, 0.0f);
    float rem = __fmaf_rn(q0, R30, x);
    return __fmaf_rn(rr, rem, q0);
}

__global__ void k_divcheck(unsigned long long* mismatches, unsigned int lo_bits, unsigned int hi_bits, float* rr_out) {
    unsigned long long local = 0;
    for (unsigned long long b = lo_bits + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; b < hi_bits;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        float x = __uint_as_float((unsigned int)b);
        volatile float dneg = -0.010000000707805156708f;
        float ref = __fdiv_rn(x, dneg);
        float fast = div_const_fast(x);
        if (__float_as_uint(ref) != __float_as_uint(fast)) local++;
    }
    if (local) atomicAdd(mismatches, local);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float R33 = 99.99999237060546875f, R30 = 0.010000000707805156708f;
        float t = __fmaf_rn(-R33, R30, 1.0f);
        rr_out[0] = __fmaf_rn(t, -R33, -R33);
    }
}

__global__ void k_unorm(float* a, float* b, float* c) {
    int k = threadIdx.x;
    a[k] = (float)k / 255.0f;
    b[k] = (float)k * (1.0f / 255.0f);
    c[k] = (float)k * 0.0039215688593685626984f;  // 0x3b808081
}

__global__ void k_ex2(const float* in, float* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __expf(in[i]);
}

int main() {
    FILE* f = fopen("gpurun_out/probe_hw.json", "w");
    if (!f) { perror("open"); return 1; }
    fprintf(f, "{\n");
    // 1. gaussian tables
    const float sig[3] = {0.5f, 1.0f, 2.0f};
    const int rad[3] = {2, 3, 6};
    float* d; cudaMalloc(&d, 256 * sizeof(float));
    for (int s = 0; s < 3; s++) {
        float s2 = sig[s] * sig[s] * 2;
        int n = 2 * rad[s] + 1;
        k_gauss<<<1, 256>>>(d, s2, rad[s]);
        std::vector<float> h(n * n);
        cudaMemcpy(h.data(), d, n * n * sizeof(float), cudaMemcpyDeviceToHost);
        fprintf(f, " \"gauss_r%d\": [", rad[s]);
        for (int i = 0; i < n * n; i++) { uint32_t u; memcpy(&u, &h[i], 4); fprintf(f, "%s%u", i ? "," : "", u); }
        fprintf(f, "],\n");
    }
    // 2. division check over x in [2^-20, 4)
    unsigned long long* dm; cudaMalloc(&dm, 8); cudaMemset(dm, 0, 8);
    float* drr; cudaMalloc(&drr, 4);
    float lo = ldexpf(1.f, -20), hi = 4.f;
    uint32_t lob, hib; memcpy(&lob, &lo, 4); memcpy(&hib, &hi, 4);
    k_divcheck<<<148 * 8, 256>>>(dm, lob, hib, drr);
    unsigned long long mm; cudaMemcpy(&mm, dm, 8, cudaMemcpyDeviceToHost);
    float rr; cudaMemcpy(&rr, drr, 4, cudaMemcpyDeviceToHost);
    uint32_t rru; memcpy(&rru, &rr, 4);
    fprintf(f, " \"div_const_mismatches\": %llu, \"div_const_range_bits\": [%u,%u], \"div_recip_bits\": %u,\n", mm, lob, hib, rru);
    // 3. unorm variants
    float *da, *db, *dc; cudaMalloc(&da, 1024); cudaMalloc(&db, 1024); cudaMalloc(&dc, 1024);
    k_unorm<<<1, 256>>>(da, db, dc);
    float ha[256], hb[256], hc[256];
    cudaMemcpy(ha, da, 1024, cudaMemcpyDeviceToHost); cudaMemcpy(hb, db, 1024, cudaMemcpyDeviceToHost); cudaMemcpy(hc, dc, 1024, cudaMemcpyDeviceToHost);
    int nb = 0, nc = 0;
    for (int k = 0; k < 256; k++) { nb += memcmp(&ha[k], &hb[k], 4) != 0; nc += memcmp(&ha[k], &hc[k], 4) != 0; }
    fprintf(f, " \"unorm_div255\": [");
    for (int k = 0; k < 256; k++) { uint32_t u; memcpy(&u, &ha[k], 4); fprintf(f, "%s%u", k ? "," : "", u); }
    fprintf(f, "],\n \"unorm_mul_recip_mismatch\": %d, \"unorm_mul_const_mismatch\": %d,\n", nb, nc);
    // 4. ex2 sample
    const int N = 4096;
    std::vector<float> xin(N), xout(N);
    for (int i = 0; i < N; i++) xin[i] = -(float)i * (100.0f / N) * (1.0f + 1e-3f * (i % 7));
    float *dxi, *dxo; cudaMalloc(&dxi, N * 4); cudaMalloc(&dxo, N * 4);
    cudaMemcpy(dxi, xin.data(), N * 4, cudaMemcpyHostToDevice);
    k_ex2<<<N / 256, 256>>>(dxi, dxo, N);
    cudaMemcpy(xout.data(), dxo, N * 4, cudaMemcpyDeviceToHost);
    double maxrel = 0; int ndiff = 0;
    for (int i = 0; i < N; i++) {
        float ref = expf(xin[i]);
        if (ref > 1e-30f) { double rel = fabs((double)xout[i] - ref) / ref; if (rel > maxrel) maxrel = rel; }
        ndiff += memcmp(&ref, &xout[i], 4) != 0;
    }
    fprintf(f, " \"expf_vs_host_maxrel\": %.3e, \"expf_vs_host_ndiff_of_4096\": %d,\n", maxrel, ndiff);
    cudaError_t e = cudaDeviceSynchronize();
    fprintf(f, " \"cuda_error\": %d\n}\n", (int)e);
    fclose(f);
    printf("probe_hw done, div mismatches=%llu unorm mism=%d/%d\n", mm, nb, nc);
    return 0;
}
