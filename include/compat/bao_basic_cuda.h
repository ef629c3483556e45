// Minimal stand-in for the reference's basic/bao_basic_cuda.h: only what callers of the class need
// (CUDA vector types and the bao_timer_gpu_cpu wall-clock helper main.cpp:60-66 uses).
#ifndef _EPPM_COMPAT_BAO_BASIC_CUDA_H_
#define _EPPM_COMPAT_BAO_BASIC_CUDA_H_
#include <cuda_runtime.h>
#include <stdio.h>
#include <chrono>

class bao_timer_gpu_cpu  // device-synchronised wall clock (basic/bao_basic_cuda.cpp:78-122)
{
public:
    void start() { cudaDeviceSynchronize(); m_t0 = std::chrono::steady_clock::now(); }
    double stop() { cudaDeviceSynchronize(); return std::chrono::duration<double>(std::chrono::steady_clock::now() - m_t0).count(); }
    double time_display(const char* disp = "", int nr_frame = 1) {
        double sec = stop() / nr_frame;
        printf("Running time (%s) is: %5.5f Seconds.\n", disp, sec);
        return sec;
    }
    double fps_display(const char* disp = "", int nr_frame = 1) {
        double fps = nr_frame / stop();
        printf("Running time (%s) is: %5.5f frame per second.\n", disp, fps);
        return fps;
    }
private:
    std::chrono::steady_clock::time_point m_t0;
};
#endif
