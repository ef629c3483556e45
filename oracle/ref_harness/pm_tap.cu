// TEST INFRASTRUCTURE ONLY — compiled into oracle/_ref/libeppm_ref.so, never into the product.
//
// The reference keeps its PatchMatch textures and RNG state in file-scope objects of
// bao_pmflow_kernel.cu (:36-39, :48), so the only way to observe the NNF between the
// kernels that baoCudaPatchMatch (:1760-1826) launches is to be in the same translation
// unit.  This file therefore #includes the reference source WHERE IT LIES (nothing is
// copied) and adds a tap that replays baoCudaPatchMatch's own launch sequence but stops
// after `n_steps` launches-groups:
//   step 1           baoGenerateRandomField   (d_setup_randgen + d_gen_rand_field)
//   step 2           baoComputeCostField
//   step 3+5*i+0..3  d_row_propagate_seg, d_column_propagate_seg,
//                    d_row_propagate_reverse_seg, d_column_propagate_reverse_seg   (iteration i)
//   step 3+5*i+4     baoRandomSearch                                               (iteration i)
#include "bao_pmflow_kernel.cu"

static void ref_bind_pm_textures(uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1,
                                 unsigned char* d_census2, int w, int h, size_t img_pitch,
                                 size_t census_pitch) {
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
    checkCudaErrors(cudaBindTexture2D(0, rgbaImg1Tex, d_img1, desc, w, h, img_pitch));
    checkCudaErrors(cudaBindTexture2D(0, rgbaImg2Tex, d_img2, desc, w, h, img_pitch));
    census1Tex.filterMode = cudaFilterModePoint;
    census1Tex.normalized = false;
    census2Tex.filterMode = cudaFilterModePoint;
    census2Tex.normalized = false;
    cudaChannelFormatDesc desc_census = cudaCreateChannelDesc<unsigned char>();
    checkCudaErrors(cudaBindTexture2D(0, census1Tex, d_census1, desc_census, w, h, census_pitch));
    checkCudaErrors(cudaBindTexture2D(0, census2Tex, d_census2, desc_census, w, h, census_pitch));
}

extern "C" int ref_tap_patchmatch(short2* d_disp_vec, float* d_cost, uchar4* d_img1, uchar4* d_img2,
                                  unsigned char* d_census1, unsigned char* d_census2, int w, int h,
                                  size_t img_pitch, size_t cost_pitch, size_t disp_pitch,
                                  size_t census_pitch, int n_steps) {
    dim3 gridSize(bao_div_ceil(w, BLOCK_DIM_X), bao_div_ceil(h, BLOCK_DIM_Y));
    dim3 blockSize(BLOCK_DIM_X, BLOCK_DIM_Y);
    checkCudaErrors(cudaMalloc(&g_d_rand_states, gridSize.x * gridSize.y * sizeof(curandState)));
    ref_bind_pm_textures(d_img1, d_img2, d_census1, d_census2, w, h, img_pitch, census_pitch);
    size_t cost_mem_w = cost_pitch / sizeof(float);
    size_t disp_mem_w = disp_pitch / sizeof(short2);

    int num_row_seg = bao_div_ceil(w, PROP_SEG_LENGTH);
    int num_col_seg = bao_div_ceil(h, PROP_SEG_LENGTH);
    dim3 gridRow(bao_div_ceil(h, ROW_PROP_SEG_BLOCK_DIM_X), bao_div_ceil(num_row_seg, ROW_PROP_SEG_BLOCK_DIM_Y));
    dim3 gridCol(bao_div_ceil(w, COL_PROP_SEG_BLOCK_DIM_X), bao_div_ceil(num_col_seg, COL_PROP_SEG_BLOCK_DIM_Y));
    dim3 blockRow(ROW_PROP_SEG_BLOCK_DIM_X, ROW_PROP_SEG_BLOCK_DIM_Y);
    dim3 blockCol(COL_PROP_SEG_BLOCK_DIM_X, COL_PROP_SEG_BLOCK_DIM_Y);

    int step = 0;
    if (step++ < n_steps) baoGenerateRandomField(d_disp_vec, w, h, disp_mem_w);
    if (step++ < n_steps) baoComputeCostField(d_cost, d_disp_vec, w, h, cost_mem_w, disp_mem_w);
    else _initGaussianLookupTable();
    for (int i = 0; i < NUM_ITER; i++) {
        if (step++ < n_steps) d_row_propagate_seg<<<gridRow, blockRow>>>(d_cost, d_disp_vec, w, h, cost_mem_w, disp_mem_w);
        if (step++ < n_steps) d_column_propagate_seg<<<gridCol, blockCol>>>(d_cost, d_disp_vec, w, h, cost_mem_w, disp_mem_w);
        if (step++ < n_steps) d_row_propagate_reverse_seg<<<gridRow, blockRow>>>(d_cost, d_disp_vec, w, h, cost_mem_w, disp_mem_w);
        if (step++ < n_steps) d_column_propagate_reverse_seg<<<gridCol, blockCol>>>(d_cost, d_disp_vec, w, h, cost_mem_w, disp_mem_w);
        if (step++ < n_steps) baoRandomSearch(d_cost, d_disp_vec, w, h, cost_mem_w, disp_mem_w);
    }
    cudaError_t e = cudaDeviceSynchronize();
    checkCudaErrors(cudaFree(g_d_rand_states));
    return (int)e;
}

// One single PatchMatch kernel applied to a caller-supplied (NNF, cost) state:
// kind 0..3 = the four segment propagations in launch order, 4 = random search preceded by
// `n_prior_searches` earlier searches (so the XORWOW states are where iteration n would find
// them: 512 draws for the initial field + 3072 per earlier search), 5 = cost field of the given NNF.
extern "C" int ref_tap_pm_step(int kind, int n_prior_searches, short2* d_disp_vec, float* d_cost,
                               uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1,
                               unsigned char* d_census2, int w, int h, size_t img_pitch,
                               size_t cost_pitch, size_t disp_pitch, size_t census_pitch) {
    dim3 gridSize(bao_div_ceil(w, BLOCK_DIM_X), bao_div_ceil(h, BLOCK_DIM_Y));
    dim3 blockSize(BLOCK_DIM_X, BLOCK_DIM_Y);
    ref_bind_pm_textures(d_img1, d_img2, d_census1, d_census2, w, h, img_pitch, census_pitch);
    _initGaussianLookupTable();
    size_t cost_mem_w = cost_pitch / sizeof(float);
    size_t disp_mem_w = disp_pitch / sizeof(short2);
    int num_row_seg = bao_div_ceil(w, PROP_SEG_LENGTH);
    int num_col_seg = bao_div_ceil(h, PROP_SEG_LENGTH);
    dim3 gridRow(bao_div_ceil(h, ROW_PROP_SEG_BLOCK_DIM_X), bao_div_ceil(num_row_seg, ROW_PROP_SEG_BLOCK_DIM_Y));
    dim3 gridCol(bao_div_ceil(w, COL_PROP_SEG_BLOCK_DIM_X), bao_div_ceil(num_col_seg, COL_PROP_SEG_BLOCK_DIM_Y));
    dim3 blockRow(ROW_PROP_SEG_BLOCK_DIM_X, ROW_PROP_SEG_BLOCK_DIM_Y);
    dim3 blockCol(COL_PROP_SEG_BLOCK_DIM_X, COL_PROP_SEG_BLOCK_DIM_Y);
    switch (kind) {
    case 0: d_row_propagate_seg<<<gridRow, blockRow>>>(d_cost, d_disp_vec, w, h, cost_mem_w, disp_mem_w); break;
    case 1: d_column_propagate_seg<<<gridCol, blockCol>>>(d_cost, d_disp_vec, w, h, cost_mem_w, disp_mem_w); break;
    case 2: d_row_propagate_reverse_seg<<<gridRow, blockRow>>>(d_cost, d_disp_vec, w, h, cost_mem_w, disp_mem_w); break;
    case 3: d_column_propagate_reverse_seg<<<gridCol, blockCol>>>(d_cost, d_disp_vec, w, h, cost_mem_w, disp_mem_w); break;
    case 4: {
        checkCudaErrors(cudaMalloc(&g_d_rand_states, gridSize.x * gridSize.y * sizeof(curandState)));
        short2* scratch_nnf; float* scratch_cost;
        checkCudaErrors(cudaMalloc(&scratch_nnf, sizeof(short2) * w * h));
        checkCudaErrors(cudaMalloc(&scratch_cost, sizeof(float) * w * h));
        baoGenerateRandomField(scratch_nnf, w, h, w);
        checkCudaErrors(cudaMemset(scratch_cost, 0, sizeof(float) * w * h));
        for (int i = 0; i < n_prior_searches; i++) baoRandomSearch(scratch_cost, scratch_nnf, w, h, w, w);
        baoRandomSearch(d_cost, d_disp_vec, w, h, cost_mem_w, disp_mem_w);
        cudaDeviceSynchronize();
        cudaFree(scratch_nnf); cudaFree(scratch_cost);
        checkCudaErrors(cudaFree(g_d_rand_states));
        break;
    }
    case 5: d_compute_cost_field<<<gridSize, blockSize>>>(d_cost, d_disp_vec, w, h, cost_mem_w, disp_mem_w); break;
    default: return -1;
    }
    return (int)cudaDeviceSynchronize();
}

// Plane-fitting refine alone on a caller-supplied (already upsampled, dense) flow plane:
// baoCudaBLFCostFilterRefine (bao_pmflow_kernel.cu:2042-2069) needs the LUTs that only
// baoComputeCostField uploads (:670-687), so upload them first like compute_flow's order does.
extern "C" int ref_tap_c2f_refine(float2* d_flow, uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1,
                                  unsigned char* d_census2, int w, int h, size_t img_pitch, size_t census_pitch) {
    _initGaussianLookupTable();
    baoCudaBLFCostFilterRefine(d_flow, d_img1, d_img2, d_census1, d_census2, w, h, img_pitch, census_pitch);
    return (int)cudaDeviceSynchronize();
}

// Device probes of hardware-defined behaviour the product must reproduce:
//  (a) unorm8 -> float conversion of cudaReadModeNormalizedFloat for all 256 values,
//  (b) which texel a point-sampled fetch at a fractional coordinate selects.
__global__ void ref_probe_unorm_kernel(float* out) {
    int k = threadIdx.x;
    float4 v = tex2D(rgbaImg1Tex, (float)k, 0.f);
    out[k] = v.x;
}
__global__ void ref_probe_point_kernel(const float* xs, int n, unsigned char* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex2D(census1Tex, xs[i], 0.f);
}
extern "C" int ref_probe_texture(float* h_unorm256, const float* h_xs, int n, unsigned char* h_texel) {
    // 256x1 uchar4 ramp (x = value) and 256x1 u8 ramp (texel value = its own index)
    uchar4 h_ramp[256]; unsigned char h_idx[256];
    for (int i = 0; i < 256; i++) { h_ramp[i] = make_uchar4(i, i, i, 0); h_idx[i] = (unsigned char)i; }
    uchar4* d_ramp; unsigned char* d_idx; size_t p4, p1;
    checkCudaErrors(cudaMallocPitch(&d_ramp, &p4, 256 * sizeof(uchar4), 1));
    checkCudaErrors(cudaMallocPitch(&d_idx, &p1, 256, 1));
    checkCudaErrors(cudaMemcpy(d_ramp, h_ramp, sizeof(h_ramp), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemcpy(d_idx, h_idx, sizeof(h_idx), cudaMemcpyHostToDevice));
    ref_bind_pm_textures(d_ramp, d_ramp, d_idx, d_idx, 256, 1, p4, p1);
    float* d_out; float* d_xs; unsigned char* d_tx;
    checkCudaErrors(cudaMalloc(&d_out, 256 * sizeof(float)));
    checkCudaErrors(cudaMalloc(&d_xs, n * sizeof(float)));
    checkCudaErrors(cudaMalloc(&d_tx, n));
    checkCudaErrors(cudaMemcpy(d_xs, h_xs, n * sizeof(float), cudaMemcpyHostToDevice));
    ref_probe_unorm_kernel<<<1, 256>>>(d_out);
    ref_probe_point_kernel<<<(n + 255) / 256, 256>>>(d_xs, n, d_tx);
    cudaError_t e = cudaDeviceSynchronize();
    checkCudaErrors(cudaMemcpy(h_unorm256, d_out, 256 * sizeof(float), cudaMemcpyDeviceToHost));
    checkCudaErrors(cudaMemcpy(h_texel, d_tx, n, cudaMemcpyDeviceToHost));
    cudaFree(d_ramp); cudaFree(d_idx); cudaFree(d_out); cudaFree(d_xs); cudaFree(d_tx);
    return (int)e;
}
