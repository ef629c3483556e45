// Microbenchmark (development tool): scattered per-lane gathers of one pixel, LSU (LDG.128 from a float4 plane) vs the texture
// unit (tex2D<float4> on a uchar4 pitch-2D texture, point sampled, unnormalised coordinates).  Pattern: every lane walks a
// 10x10 stride-2 window around its own random centre, like the target side of a PatchMatch random guess.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void k_ldg(const float4* __restrict__ plane, int pw, const short2* __restrict__ centres, float* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    short2 c = centres[i];
    float acc = 0.f;
    for (int dy = -9; dy <= 9; dy += 2)
#pragma unroll
        for (int dx = -9; dx <= 9; dx += 2) {
            float4 p = __ldg(plane + (c.y + 16 + dy) * pw + c.x + 16 + dx);
            acc += p.x + p.y * 0.5f + p.z * 0.25f + __uint_as_float(__float_as_uint(p.w) & 0x3f800000u);
        }
    out[i] = acc;
}
__global__ void k_tex(cudaTextureObject_t tex, const short2* __restrict__ centres, float* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    short2 c = centres[i];
    float acc = 0.f;
    const float fx = (float)c.x, fy = (float)c.y;
    for (int dy = -9; dy <= 9; dy += 2)
#pragma unroll
        for (int dx = -9; dx <= 9; dx += 2) {
            float4 p = tex2D<float4>(tex, fx + (float)dx, fy + (float)dy);
            acc += p.x + p.y * 0.5f + p.z * 0.25f + p.w;
        }
    out[i] = acc;
}

__global__ void k_mix(const float4* __restrict__ plane, int pw, cudaTextureObject_t tex, const short2* __restrict__ centres, float* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    short2 c = centres[i];
    float acc = 0.f;
    const float fx = (float)c.x, fy = (float)c.y;
    for (int dy = -9; dy <= 9; dy += 2)
#pragma unroll
        for (int dx = -9; dx <= 9; dx += 4) {   // alternate: one sample through the LSU, the next through the texture unit
            float4 p = __ldg(plane + (c.y + 16 + dy) * pw + c.x + 16 + dx);
            acc += p.x + p.y * 0.5f + p.z * 0.25f + __uint_as_float(__float_as_uint(p.w) & 0x3f800000u);
            float4 q = tex2D<float4>(tex, fx + (float)(dx + 2), fy + (float)dy);
            acc += q.x + q.y * 0.5f + q.z * 0.25f + q.w;
        }
    out[i] = acc;
}

int main(int argc, char** argv) {
    const int w = 480, h = 270, pw = w + 32, ph = h + 32, n = 480 * 270 * 16;
    const int coherent = argc > 1 ? atoi(argv[1]) : 0;
    std::vector<float4> plane((size_t)pw * ph);
    std::vector<uchar4> img((size_t)w * h);
    for (auto& p : plane) p = make_float4(rand() % 256 / 255.f, rand() % 256 / 255.f, rand() % 256 / 255.f, 0.f);
    for (auto& p : img) p = make_uchar4(rand() % 256, rand() % 256, rand() % 256, rand() % 256);
    std::vector<short2> cen(n);
    for (int i = 0; i < n; i++) {
        int x = i % w, y = (i / w) % h;
        if (coherent) cen[i] = make_short2((x + 5) % w, (y + 3) % h);          // neighbouring lanes -> neighbouring targets
        else cen[i] = make_short2(rand() % w, rand() % h);                     // random guess
    }
    float4* d_plane; uchar4* d_img; short2* d_cen; float* d_out; size_t pitch;
    CK(cudaMalloc(&d_plane, plane.size() * 16)); CK(cudaMallocPitch(&d_img, &pitch, w * 4, h));
    CK(cudaMalloc(&d_cen, n * 4)); CK(cudaMalloc(&d_out, n * 4));
    CK(cudaMemcpy(d_plane, plane.data(), plane.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy2D(d_img, pitch, img.data(), w * 4, w * 4, h, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_cen, cen.data(), n * 4, cudaMemcpyHostToDevice));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypePitch2D; rd.res.pitch2D.devPtr = d_img; rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
    rd.res.pitch2D.width = w; rd.res.pitch2D.height = h; rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, NULL));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 3; which++) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            if (which == 0) k_ldg<<<(n + 127) / 128, 128>>>(d_plane, pw, d_cen, d_out, n);
            else if (which == 1) k_tex<<<(n + 127) / 128, 128>>>(tex, d_cen, d_out, n);
            else k_mix<<<(n + 127) / 128, 128>>>(d_plane, pw, tex, d_cen, d_out, n);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%s %s: %.3f ms for %d lanes x 100 gathers = %.1f Ggather/s\n", coherent ? "coherent" : "random", which == 0 ? "LDG.128 float4" : which == 1 ? "TEX uchar4" : "half LDG half TEX", best, n,
               n * 100.0 / best / 1e6);
    }
    return 0;
}
