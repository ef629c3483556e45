/* TEST INFRASTRUCTURE ONLY.  Device-defined constants pinned on a B200 (tools/probe_hw.cu, gpurun 2026-10-17):
 * bit patterns of __expf(-(float)(dy*dy+dx*dx)/(2*sigma*sigma)) for the two pyramid blurs of the reference
 * (basic/bao_basic_cuda.cuh:452; sigma .5 radius 2 = pre-blur, sigma 1 radius 3 = every pyramid level), row-major
 * (dy outer, dx inner).  MUFU.EX2 cannot be reproduced on a CPU, so the oracle uses the GPU's own values: with them
 * the blur (fmaf accumulation + IEEE division + truncation) is bit-exact on the host.
 * Also probed there: cudaReadModeNormalizedFloat returns exactly RN(k/255.f) for all 256 k; point-sampled fetches at
 * fractional coordinates select floor(x) (no fixed-point rounding); x / -(0.1f*0.1f) via the 3-FMA fast path equals
 * div.rn for every float in [2^-20, 4). */
#include <stdint.h>
static const uint32_t kGaussR2Bits[25] = {
    0x33f1aae1u, 0x383e6bcdu, 0x39afe109u, 0x383e6bcdu, 0x33f1aae1u, 0x383e6bcdu, 0x3c960aaeu,
    0x3e0a9555u, 0x3c960aaeu, 0x383e6bcdu, 0x39afe109u, 0x3e0a9555u, 0x3f800000u, 0x3e0a9555u,
    0x39afe109u, 0x383e6bcdu, 0x3c960aaeu, 0x3e0a9555u, 0x3c960aaeu, 0x383e6bcdu, 0x33f1aae1u,
    0x383e6bcdu, 0x39afe109u, 0x383e6bcdu, 0x33f1aae1u};
static const uint32_t kGaussR3Bits[49] = {
    0x39016794u, 0x3ac50f0cu, 0x3bdcc9feu, 0x3c360284u, 0x3bdcc9feu, 0x3ac50f0cu, 0x39016794u,
    0x3ac50f0cu, 0x3c960aaeu, 0x3da81c2du, 0x3e0a9555u, 0x3da81c2du, 0x3c960aaeu, 0x3ac50f0cu,
    0x3bdcc9feu, 0x3da81c2du, 0x3ebc5ab2u, 0x3f1b4598u, 0x3ebc5ab2u, 0x3da81c2du, 0x3bdcc9feu,
    0x3c360284u, 0x3e0a9555u, 0x3f1b4598u, 0x3f800000u, 0x3f1b4598u, 0x3e0a9555u, 0x3c360284u,
    0x3bdcc9feu, 0x3da81c2du, 0x3ebc5ab2u, 0x3f1b4598u, 0x3ebc5ab2u, 0x3da81c2du, 0x3bdcc9feu,
    0x3ac50f0cu, 0x3c960aaeu, 0x3da81c2du, 0x3e0a9555u, 0x3da81c2du, 0x3c960aaeu, 0x3ac50f0cu,
    0x39016794u, 0x3ac50f0cu, 0x3bdcc9feu, 0x3c360284u, 0x3bdcc9feu, 0x3ac50f0cu, 0x39016794u};
