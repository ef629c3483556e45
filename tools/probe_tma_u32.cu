// Development probe: which tiled tensor maps over a dense uchar4 plane does cp.async.bulk.tensor accept?  (element type x start coordinate)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build/probe_tma_u32 tools/probe_tma_u32.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int bytes, unsigned* out) {
    __shared__ __align__(128) unsigned tile[2048];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(tile)),
                     "l"(&tmap), "r"(c0), "r"(c1), "r"(0), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    if (threadIdx.x < 4) out[threadIdx.x] = tile[threadIdx.x];
}
int main() {
    const int w = 128, h = 96;
    unsigned* d; cudaMalloc(&d, w * h * 4);
    unsigned* hbuf = (unsigned*)malloc(w * h * 4);
    for (int i = 0; i < w * h; i++) hbuf[i] = i;
    cudaMemcpy(d, hbuf, w * h * 4, cudaMemcpyHostToDevice);
    unsigned* out; cudaMalloc(&out, 16);
    struct { CUtensorMapDataType t; int es; const char* n; } types[] = {{CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "u32"}, {CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, "f32"}, {CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, "u8"}};
    for (auto& ty : types)
        for (int bw : {36, 48}) {
            CUtensorMap m;
            cuuint64_t dims[3] = {(cuuint64_t)w * 4 / ty.es, (cuuint64_t)h, 1}, strides[2] = {(cuuint64_t)w * 4, (cuuint64_t)w * h * 4};
            cuuint32_t box[3] = {(cuuint32_t)(bw * 4 / ty.es), 10, 1}, es[3] = {1, 1, 1};
            CUresult r = cuTensorMapEncodeTiled(&m, ty.t, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            for (int c0px : {16, 15}) {
                cudaMemset(out, 0xff, 16);
                k<<<1, 32>>>(m, c0px * 4 / ty.es, 7, bw * 4 * 10, out);
                cudaError_t e = cudaDeviceSynchronize();
                unsigned o[4] = {0, 0, 0, 0};
                if (e == cudaSuccess) cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost);
                printf("type %s box %d px start %d px: encode %d run %s first %u (expect %d)\n", ty.n, bw, c0px, (int)r, cudaGetErrorString(e), o[0], 7 * w + c0px);
                if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 0; }
            }
        }
    return 0;
}
