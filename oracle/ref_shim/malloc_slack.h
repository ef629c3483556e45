// TEST INFRASTRUCTURE ONLY — force-included into every translation unit of the oracle/_ref build.
//
// The reference's pyramid build writes the blurred level-1 image into a DENSE scratch buffer
// (bao_cuda_pyr_alloc<uchar4>, bao_flow_patchmatch_multiscale_cuda.cpp:155-156) using the PITCHED row stride of the
// image pyramid (bao_cuda_gauss_filter_pitched(pPyrTemp[i-n], ..., arrPitch[i-n], ...), basic/bao_basic_cuda.cuh:660).
// Whenever (w/2)*4 is not a multiple of the cudaMallocPitch alignment the kernel writes h/2*(pitch - w/2*4) bytes
// past the end of that buffer: 138 KB at 1920x1080, where compute-sanitizer pins it to bao_basic_cuda.cuh:466 and the
// unmodified build dies with cudaErrorIllegalAddress on B200.  To obtain a baseline at all, every cudaMalloc of the
// reference build gets 1 MiB of slack so the stray rows land in memory the same buffer owns.  No reference source is
// changed and the values the reference computes are unaffected (the stray rows are read back with the same stride).
#pragma once
#include <cuda_runtime.h>
#ifndef REF_MALLOC_SLACK
#define REF_MALLOC_SLACK ((size_t)1 << 20)
#endif
#define cudaMalloc(p, s) cudaMalloc((p), (size_t)(s) + REF_MALLOC_SLACK)
