/* eppm_legacy_abi.h — the reference's own stage functions, re-exported by libeppm_b200.so with the reference's
 * signatures and buffer layouts, so that linchaobao/EPPM's host class (bao_flow_patchmatch_multiscale_cuda.cpp)
 * links against this library unchanged.  Declarations follow bao_flow_patchmatch_multiscale_cuda.cpp:40-62.
 *
 * All pointers are DEVICE pointers; pyramids are host arrays of device pointers (basic/bao_basic_cuda.h:209-229);
 * pitches are in bytes; images are uchar4 (alpha ignored) and census planes u8, both possibly pitched;
 * NNF = short2 absolute target (x,y), cost = float, flow = float2.  Every function returns void like the reference:
 * CUDA errors are printed to stderr (helper_cuda.h:682-712 behaviour) and can be read with eppm_last_error().
 * The functions synchronise with the legacy default stream on entry and are complete on return.
 */
#ifndef EPPM_LEGACY_ABI_H_
#define EPPM_LEGACY_ABI_H_
#include <stddef.h>
#include <cuda_runtime.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bao_pmflow_refine_kernel.cu:1060-1071 */
void baoCudaPatchMatchMultiscalePrepare(uchar4** pImgPyr1, uchar4** pImgPyr2, unsigned char** pCensusPyr1, unsigned char** pCensusPyr2,
                                        uchar4** pTempPyr1, uchar4** pTempPyr2, int* arrH, int* arrW, size_t* arrPitchUchar4,
                                        size_t* arrPitchUchar1, int nLevels, uchar4* d_img1, uchar4* d_img2, int h, int w);
/* bao_pmflow_census_kernel.cu:93-112 */
void baoCudaCensusTransform(unsigned char* d_census1, unsigned char* d_census2, uchar4* d_img1, uchar4* d_img2, int w, int h, size_t img_pitch,
                            size_t census_pitch);
/* bao_pmflow_kernel.cu:1760-1826 */
void baoCudaPatchMatch(short2* d_disp_vec, float* d_cost, uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1, unsigned char* d_census2,
                       int w, int h, size_t img_pitch, size_t cost_pitch, size_t disp_pitch, size_t census_pitch);
/* bao_pmflow_refine_kernel.cu:78-92 */
void baoCudaLeftRightCheck(short2* d_disp_vec, float* d_cost, short2* d_disp_vec2, float* d_cost2, int w, int h, size_t cost_pitch,
                           size_t disp_pitch);
/* :185-193 */
void baoCudaOutlierRemoval(short2* d_disp_vec, float* d_cost, int w, int h, size_t cost_pitch, size_t disp_pitch);
/* :261-286 */
void baoCudaWeightedMedianFilter(short2* d_disp_vec, float* d_cost, uchar4* d_img, int w, int h, size_t img_pitch, size_t cost_pitch,
                                 size_t disp_pitch, int num_iter, bool is_only_occlusion);
/* :373-390 */
void baoCudaFillHole(short2* d_disp_vec, float* d_cost, uchar4* d_img, int w, int h, size_t img_pitch, size_t cost_pitch, size_t disp_pitch);
/* :724-734 */
void baoCudaNNF2Flow(float2* d_flow, short2* d_disp_vec, int w, int h, size_t disp_pitch, size_t flow_pitch);
/* :1076-1087 */
void baoCudaBLF_C2F(float2** pFlowPyr, uchar4** pImgPyr1, uchar4** pImgPyr2, unsigned char** pCensusPyr1, unsigned char** pCensusPyr2,
                    float2** pTempPyr1, float2** pTempPyr2, int* arrH, int* arrW, size_t* arrPitchUchar4, size_t* arrPitchUchar1, int nLayerIdx);
/* bao_pmflow_kernel.cu:2042-2069 (plane-fitting refine alone, on an already upsampled dense flow plane) */
void baoCudaBLFCostFilterRefine(float2* d_flow_vec, uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1, unsigned char* d_census2, int w,
                                int h, size_t img_pitch, size_t census_pitch);
/* bao_pmflow_refine_kernel.cu:801-826 */
void baoCudaFlowSmoothing(float2* d_flow, uchar4* d_img, int w, int h, size_t img_pitch, size_t flow_pitch);

/* ---- declared by the reference's host class (…cuda.cpp:40-62) but not called by compute_flow ---- */
/* bao_pmflow_refine_kernel.cu:93-140: left-right check with threshold 50 through caller-provided temp planes (dense h*w) */
void baoCudaLeftRightCheck_Buffered(short2* d_disp_vec, float* d_cost, short2* d_disp_vec2, float* d_cost2, short2* d_disp_vec_temp,
                                    float* d_cost_temp, int w, int h, size_t cost_pitch, size_t disp_pitch);
/* bao_pmflow_refine_kernel.cu:657-676, 736-746: flow -> absolute targets, unknown flow -> INVALID_LOCATION */
void baoCudaFlow2NNF(short2* d_disp_vec, float2* d_flow, int w, int h, size_t disp_pitch, size_t flow_pitch);
/* bao_pmflow_refine_kernel.cu:891-912: clamp both components to [-max_flow_val, max_flow_val] */
void baoCudaFlowCutoff(float2* d_flow, int w, int h, size_t flow_pitch, float max_flow_val);
/* bao_pmflow_kernel.cu:555-586, 2071-2095: zero the flow where the two images already agree (mean AD term of the patch <= 0.1); dense flow plane */
void baoEliminateStillRegionFlow(float2* d_flow, uchar4* d_img1, uchar4* d_img2, int w, int h, size_t img_pitch);
/* bao_pmflow_refine_kernel.cu:976-1057: the guide image filtered with its own joint-bilateral weights (the reference's preceding 5x5 median
 * is overwritten by it and not executed here); alpha of the output is 0 (uninitialised in the reference) */
void baoCudaImageSmoothing(uchar4* d_img_smoothed, uchar4* d_img, int w, int h, size_t img_pitch);
/* bao_pmflow_refine_kernel.cu:829-888: joint-bilateral upsampling of a coarser flow (dense planes), values scaled by ratio_up; pixels
 * without any known tap keep their content (the spelling of the name is the reference's) */
void baoCudaFlowBilteralUpsampling(float2* d_flow_vec, uchar4* d_img, int w, int h, size_t img_pitch, float2* d_flow_vec_small, int w_s, int h_s,
                                   float ratio_up);
/* bao_pmflow_kernel.cu:1828-1895: PatchMatch over (target, patch scale in [0.6, 1.4]) scored with the bilateral AD cost alone.  Unfinished
 * upstream and mirrored as it stands (its forward row pass writes the winning candidate's scale into the cost plane, :1207; the census planes
 * are never read and may be NULL here): bit-exact against the reference build.  scale_pitch must equal disp_pitch (the reference's random-field
 * kernel indexes the scale plane with the displacement pitch, :151); otherwise, or on a null plane, it reports on stderr, sets
 * eppm_last_error() and leaves the outputs untouched. */
void baoCudaPatchMatch_Scaled(short2* d_disp_vec, float* d_scale, float* d_cost, uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1,
                              unsigned char* d_census2, int w, int h, size_t img_pitch, size_t cost_pitch, size_t disp_pitch, size_t scale_pitch,
                              size_t census_pitch);
/* bao_pmflow_kernel.cu:1897-1963: PatchMatch scored with the plane-fitting cost of the refine stage (min over four affine patch models) */
void baoCudaPatchMatch_PlaneFitting(short2* d_disp_vec, float* d_cost, uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1,
                                    unsigned char* d_census2, int w, int h, size_t img_pitch, size_t cost_pitch, size_t disp_pitch, size_t census_pitch);
/* bao_pmflow_census_kernel.cu:115-181: 3x3 census of the bicubic (B-spline) upsampled images, [h_up][w_up] u8 planes */
void baoCudaCensusTransform_Bicubic(unsigned char* d_census1, unsigned char* d_census2, int w_up, int h_up, size_t census_pitch, uchar4* d_img1,
                                    uchar4* d_img2, int w, int h, size_t img_pitch);
/* bao_pmflow_refine_kernel.cu:395-634, 678-722: sub-pixel refinement of valid integer targets (5x5 half-pixel cost samples, least-squares
 * quadric, <= 5 conjugate-gradient steps); census planes are those of baoCudaCensusTransform_Bicubic at [2h][2w]; pixels whose target is
 * out of range, or whose quadric has no stationary point within 3 half-pixels, keep their flow */
void baoCudaSubpixRefine(float2* d_flow, short2* d_disp_vec, uchar4* d_img1, uchar4* d_img2, unsigned char* d_census1_up, unsigned char* d_census2_up,
                         int w, int h, size_t img_pitch, size_t census_pitch_up, size_t disp_pitch, size_t flow_pitch);

#ifdef __cplusplus
}
/* C++ linkage like the reference (basic/bao_basic_cuda.cuh:838-849; declared at …cuda.cpp:64 and called at :311 when compute_flow is
 * given a colour buffer): Middlebury colour coding of a dense flow plane on the device, unknown flow = black. */
void bao_cuda_convert_flow_to_colorshow(uchar4* rgbflow, float2* flow_vec, int h, int w, float max_disp_x = 100, float max_disp_y = 100);
#endif
#endif
