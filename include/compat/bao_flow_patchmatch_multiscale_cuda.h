// Drop-in replacement for the reference's class header (bao_flow_patchmatch_multiscale_cuda.h:33-45 of linchaobao/EPPM):
// same class name, same public members, same calling sequence, so main.cpp compiles unmodified with
//   -I include/compat   and links against  eppm_b200/libeppm_b200.so.
// State lives behind an opaque eppm_context (include/eppm.h); no host pointer is retained past a call.
#ifndef _BAO_FLOW_PATCHMATCH_MULTISCALE_CUDA_H_
#define _BAO_FLOW_PATCHMATCH_MULTISCALE_CUDA_H_

#include "bao_basic_cuda.h"

struct eppm_context;

class bao_flow_patchmatch_multiscale_cuda
{
public:
    bao_flow_patchmatch_multiscale_cuda();
    ~bao_flow_patchmatch_multiscale_cuda();

    // interface (identical to the reference)
    void init(int h, int w);
    void init(unsigned char*** img1, unsigned char*** img2, int h, int w);
    bool set_data(unsigned char*** img1, unsigned char*** img2);  // always true, like the reference (…cuda.cpp:159-168)
    void compute_flow(float** disp1_x, float** disp1_y, unsigned char*** color_flow = NULL);

private:
    bao_flow_patchmatch_multiscale_cuda(const bao_flow_patchmatch_multiscale_cuda&);
    void operator=(const bao_flow_patchmatch_multiscale_cuda&);
    eppm_context* m_ctx;
    int m_h, m_w;
    unsigned char* m_d_rgb[2];  // device copies of the current pair, packed RGB
    float* m_d_flow;            // device result, interleaved (u,v)
    float* m_h_flow;            // pinned staging for the result
    bool m_has_data;
};

#endif
