// Device-side arithmetic primitives shared by all kernels of libeppm_b200.
//
// Every floating-point step is written with explicit round-to-nearest intrinsics so that nvcc can
// neither contract nor re-associate it.  The sequences restate, operation for operation, what
// nvcc 12.9 generates for the reference's expressions at its default flags (-fmad=true, no fast-math);
// the reference file:line each one follows is given beside it.  Results are therefore bit-identical
// to the reference's kernels given identical inputs.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace eppm {

constexpr int PAD = 16;        // replicated border of the packed planes (>= 9 + plane-fitting reach 5.2, and >= 10 at the PatchMatch level)
constexpr int PATCH_R = 9;     // defs.h:42
constexpr int INVALID_LOCATION = -10000;  // bao_pmflow_refine_kernel.cu:46
#define EPPM_UNKNOWN_FLOW 1e10f            // defs.h:89-91
#define EPPM_UNKNOWN_FLOW_THRESH 1e9f      // defs.h:84-86

// Packed pixel of a pyramid level: r,g,b as the floats a cudaReadModeNormalizedFloat fetch returns
// (exactly RN(k/255), probed on B200) and the 3x3 census byte in the bits of .w.
// One 16-byte load replaces the reference's two texture fetches (colour + census) per sample.
struct PlaneRef {
    const float4* p;   // points at logical pixel (0,0) of one image inside its padded plane
    int pw;            // padded row pitch in pixels
};

__device__ __forceinline__ float4 ldpix(const float4* p) { return __ldg(p); }

// LUTs the reference uploads to __constant__ memory on every call (bao_pmflow_kernel.cu:670-687);
// here they ride in kernel parameter space (constant bank 0), so contexts never share mutable state.
struct CostLut {
    float gg[10][10];   // gg[|i|][|j|] = G[|j|]*G[|i|], G[k] = expf(-k^2/sigma_s^2)   (:293, :676)
    float census[9];    // 1 - expf(-d^2/(lambda_census*8)^2)                              (:683)
    float pad_[3];
};

// x / -(0.1f*0.1f), correctly rounded.  nvcc lowers the reference's `-(c*c)/(LAMBDA_AD*LAMBDA_AD)` and
// `-(w+t)/(PM_SIG_R*PM_SIG_R)` (bao_pmflow_kernel.cu:282,288) to a division by the folded constant
// -0.010000000707805157 whose fast path is q0=x*r, rem=fma(q0,0.01',x), q=fma(r,rem,q0) with r=-99.99999237
// (FCHK guards only denormal/huge operands).  tools/probe_hw.cu verified on B200 that this equals div.rn for
// EVERY float in [2^-20, 4); the operands here are 0 or in [1.5e-5, 2].
__device__ __forceinline__ float div_neg_0p01(float x) {
    const float r = -99.99999237060546875f;
    const float d = 0.010000000707805156708f;
    float q0 = __fmaf_rn(x, r, 0.0f);
    float rem = __fmaf_rn(q0, d, x);
    return __fmaf_rn(r, rem, q0);
}

// ex2.approx on the MUFU pipe.  The .ftz form is the bare MUFU.EX2; results in the normal range are the same bits the
// reference's non-ftz __expf path gets from the same instruction.
__device__ __forceinline__ float ex2_mufu(float t) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}

// __expf(x) exactly as nvcc lowers it without -ftz (reference SASS): t = x*log2e; if (t < -126) { t *= .5; r = ex2(t); r *= r } else r = ex2(t).
__device__ __forceinline__ float exp_ref(float x) {
    float t = __fmul_rn(x, 1.4426950216293334961f);
    const bool tiny = t < -126.0f;
    if (tiny) t = __fmul_rn(t, 0.5f);
    float r = ex2_mufu(t);
    if (tiny) r = __fmul_rn(r, r);
    return r;
}

// 1 - __expf(x) for x <= 0.  When x*log2e < -126 the reference's fix-up yields e < 2^-126, and 1 - e rounds to exactly
// 1.0f whatever e is, so the fix-up (3 issue slots) can be dropped without changing a bit: MUFU.EX2 returns 0 there.
__device__ __forceinline__ float one_minus_exp_ref(float x) {
    return __fadd_rn(1.0f, -ex2_mufu(__fmul_rn(x, 1.4426950216293334961f)));
}

// the AD term 1 - __expf(-(c*c) / (LAMBDA_AD*LAMBDA_AD)) of one sample (bao_pmflow_kernel.cu:282-283, :576), scalar form
__device__ __forceinline__ float exp_ad_cost(float c) { return one_minus_exp_ref(div_neg_0p01(__fmul_rn(c, c))); }

__device__ __forceinline__ float max3abs_diff(const float4& a, const float4& b) {
    float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
    return fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz));
}

// Census-distance LUT in shared memory, addressed by BYTE offset: the packed planes keep the census byte replicated in
// all four bytes of .w, so popc(w1 ^ w2) = 4 * hamming distance = the byte offset of the LUT entry (no mask, no shift).
// (A register-indexed constant-bank read would serialise on the up to 9 distinct indices of a warp; a 256-entry table
// indexed by the XOR itself avoids POPC but was measured slower: its bank conflicts cost more L1 wavefronts than POPC costs XU slots.)
constexpr int CENSUS_LUT_N = 9;
__device__ __forceinline__ float census_lut(const float* s_census, const float4& p1, const float4& p2) {
    const unsigned off = __popc(__float_as_uint(p1.w) ^ __float_as_uint(p2.w));
    return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(s_census) + off);
}
// The same lookup through a 32-bit shared-window address the caller computed ONCE (census_lut_base) and keeps in a register: on
// sm_100 every shared address carries the CTA's rank in its cluster, and left alone ptxas re-derives that base (S2UR SR_CgaCtaId +
// three uniform instructions, with the S2UR latency exposed) inside every block of the sample loop.
__device__ __forceinline__ unsigned census_lut_base(const float* s_census) {
    unsigned b = (unsigned)__cvta_generic_to_shared(s_census);
    asm volatile("" : "+r"(b));
    return b;
}
__device__ __forceinline__ float census_lut_at(unsigned base, const float4& p1, const float4& p2) {
#ifdef EPPM_WHATIF_NOPOPC   // timing experiment only (wrong results): how much of the kernel time is the XU pipe's POPC?
    const unsigned off = (__float_as_uint(p1.w) ^ __float_as_uint(p2.w)) & 0x1cu;
#else
    const unsigned off = __popc(__float_as_uint(p1.w) ^ __float_as_uint(p2.w));
#endif
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + off));
    return v;
}
__device__ __forceinline__ void load_census_lut(float* s_census, const CostLut& lut) {
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    if (tid < CENSUS_LUT_N) s_census[tid] = lut.census[tid];
    __syncthreads();
}
__device__ __forceinline__ unsigned pack_census(unsigned census_byte) { return census_byte * 0x01010101u; }
__device__ __forceinline__ unsigned unpack_census(float w) { return __float_as_uint(w) & 0xffu; }

// Packed FP32x2 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2): two IEEE round-to-nearest operations per issue slot.  Each
// half is rounded exactly like the scalar instruction, so packing changes the instruction count, not a single bit.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

struct PixPk {  // a packed-plane pixel viewed as two FP32 pairs: (r,g) and (b, census bits)
    f32x2 xy, zw;
};
__device__ __forceinline__ PixPk pack_pix(const float4& p) { return PixPk{pk2(p.x, p.y), pk2(p.z, p.w)}; }

// max(|a.x-b.x|, |a.y-b.y|, |a.z-b.z|): x and y as one packed subtraction, z as a scalar one (a packed instruction holds the FMA
// pipe for two cycles, and the .w half of a second packed subtraction would be thrown away)
__device__ __forceinline__ float max3abs_diff(const PixPk& a, const PixPk& b) {
    float dx, dy, az, aw, bz, bw;
    upk2(sub2(a.xy, b.xy), dx, dy);
    upk2(a.zw, az, aw);
    upk2(b.zw, bz, bw);
    return fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(__fsub_rn(az, bz)));
}

// One sample of the bilateral-weighted AD+census patch cost (bao_pmflow_kernel.cu:274-296):
//   cost   = 1 - exp(-c^2/lambda_ad^2) + LUT_census[popc(census1 ^ census2)],  c = max|rgb1-rgb2|
//   weight = exp(-(d1^2 + d2^2)/sigma_r^2) * G[|j|]*G[|i|],                     dk = max|centre_k - p_k|
// accumulated as cost_sum = fma(cost, weight, cost_sum); weight_sum += weight, in sample order.
// d1 (image-1 side) is passed in so callers can hoist it across candidates.  The AD chain
// (c^2 -> /-0.01 -> *log2e) and the weight chain (arg -> /-0.01 -> *log2e) run side by side in the two halves of packed instructions.
// sample_eval returns the AD+census cost and the exponent t2 of the range weight; the caller forms e2 = __expf-equivalent of t2
// (ex2 with the `t2 < -126` fix-up, which a group of samples can share one test for), w = e2*gg, and accumulates.
__device__ __forceinline__ float census_lut_ref(const float* s, const float4& p1, const float4& p2) { return census_lut(s, p1, p2); }
__device__ __forceinline__ float census_lut_ref(unsigned base, const float4& p1, const float4& p2) { return census_lut_at(base, p1, p2); }
// Census LUT at a FIXED address of the shared window: the kernel keeps the table at the start of its dynamic shared memory and owns no
// static shared memory, so the table sits at the window's user base (1 KB reserved by the system on sm_100, probed per device by
// lut0_window_base_ok()) and popc(w1 ^ w2) + that constant is the whole address: the immediate field of LDS replaces one integer add per sample.
constexpr unsigned SHARED_WINDOW_USER_BASE = 0x400;
struct Lut0 {};
__device__ __forceinline__ float census_lut_ref(Lut0, const float4& p1, const float4& p2) {
    const unsigned off = __popc(__float_as_uint(p1.w) ^ __float_as_uint(p2.w));
    float v;
    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(off), "n"(SHARED_WINDOW_USER_BASE));
    return v;
}
// Census cost looked up by the XOR BYTE itself (256 entries) instead of by its popcount (9 entries): no POPC, which shares the XU pipe with the
// two MUFU.EX2 of a sample (the refine kernel's top stall is the XU pipe's throttle).  A plain 256-entry table would serialise on bank
// conflicts (32 lanes, arbitrary entries), so the table is replicated REP times: entry e, copy k at word e * REP + k, lane l reads copy
// l % REP -- lanes of different copies never share a bank.  The table sits at the shared-window base (see Lut0); `lane_off` = (lane % REP) * 4.
template <int REP>
struct LutX {
    unsigned lane_off;
};
template <int REP>
__device__ __forceinline__ float census_lut_ref(LutX<REP> l, const float4& p1, const float4& p2) {
    const unsigned e = (__float_as_uint(p1.w) ^ __float_as_uint(p2.w)) & 0xffu;
    const unsigned off = e * (REP * 4) + l.lane_off;
    float v;
    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(off), "n"(SHARED_WINDOW_USER_BASE));
    return v;
}
template <class LutRef>
__device__ __forceinline__ void sample_eval(const float4& p1, const PixPk& p1k, const float4& p2, const PixPk& c2k, float d1, LutRef s_census,
                                            float& cost, float& t2) {
    const PixPk p2k = pack_pix(p2);
    const float c = max3abs_diff(p1k, p2k);
    const float d2 = max3abs_diff(c2k, p2k);
    // c^2 and arg = fma(d1, d1, d2*d2) as scalar instructions whose results land directly in the two halves of one register pair
    // (squaring (c, d2) as a pair and patching one half afterwards costs two register copies)
    const f32x2 x = pk2(__fmul_rn(c, c), __fmaf_rn(d1, d1, __fmul_rn(d2, d2)));
    const f32x2 R = pk2(-99.99999237060546875f, -99.99999237060546875f), D = pk2(0.010000000707805156708f, 0.010000000707805156708f);
    const f32x2 q0 = fma2(x, R, pk2(0.f, 0.f));
    const f32x2 rem = fma2(q0, D, x);
    const f32x2 q = fma2(R, rem, q0);
    float t1;
    upk2(mul2(q, pk2(1.4426950216293334961f, 1.4426950216293334961f)), t1, t2);
#ifdef EPPM_WHATIF_NOEX2   // timing experiment only (wrong results): the AD term without its MUFU.EX2
    cost = __fadd_rn(__fadd_rn(1.0f, -t1), census_lut_ref(s_census, p1, p2));
#else
    cost = __fadd_rn(__fadd_rn(1.0f, -ex2_mufu(t1)), census_lut_ref(s_census, p1, p2));
#endif
}
// __expf of an exponent already multiplied by log2e, for the rare t < -126 case (see exp_ref)
__device__ __forceinline__ float ex2_tiny(float t) {
    const float r = ex2_mufu(__fmul_rn(t, 0.5f));
    return __fmul_rn(r, r);
}

// TWO samples (a, b) side by side: everything behind the two max|.| reductions of a sample -- the squares, both constant divisions,
// the log2e multiplies, 1 - e, + census -- runs in the halves of packed FP32x2 instructions, one issue slot for the two samples (the
// scalar form above pairs the AD chain with the weight chain of ONE sample and leaves squares, 1 - e, + census, e*gg and the
// accumulation scalar: 29.5 issued instructions per sample in the refine kernel against ~23 here).  Every half is rounded like the
// scalar instruction, so each sample's (cost, t2) is the bit pattern sample_eval returns.
//   p1a/p1b: image-1 pixel of sample a / b (the same pixel when a and b are two models of one candidate), p2a/p2b: image-2 pixels,
//   c2a/c2b: candidate centres in image 2, d1: (d1_a, d1_b) range distances on the image-1 side.
// Returns ct = (cost_a, cost_b) and t2 = (log2e-scaled exponent of the range weight of a, of b).
template <class LutRef>
__device__ __forceinline__ void sample_eval2(const float4& p1a, const PixPk& p1ka, const float4& p1b, const PixPk& p1kb, const float4& p2a, const float4& p2b,
                                             const PixPk& c2a, const PixPk& c2b, f32x2 d1, LutRef lut_base, f32x2& ct, f32x2& t2) {
    const PixPk p2ka = pack_pix(p2a), p2kb = pack_pix(p2b);
    const f32x2 c = pk2(max3abs_diff(p1ka, p2ka), max3abs_diff(p1kb, p2kb));
    const f32x2 d2 = pk2(max3abs_diff(c2a, p2ka), max3abs_diff(c2b, p2kb));
    const f32x2 xc = mul2(c, c);                          // c^2                          (bao_pmflow_kernel.cu:282)
    const f32x2 xw = fma2(d1, d1, mul2(d2, d2));          // fma(d1, d1, d2*d2)           (:288 as contracted)
    const f32x2 R = pk2(-99.99999237060546875f, -99.99999237060546875f), D = pk2(0.010000000707805156708f, 0.010000000707805156708f);
    const f32x2 Z = pk2(0.f, 0.f), L2E = pk2(1.4426950216293334961f, 1.4426950216293334961f);
    const f32x2 qc0 = fma2(xc, R, Z), qw0 = fma2(xw, R, Z);                        // div_neg_0p01 on both pairs
    const f32x2 qc = fma2(R, fma2(qc0, D, xc), qc0), qw = fma2(R, fma2(qw0, D, xw), qw0);
    float t1a, t1b;
    upk2(mul2(qc, L2E), t1a, t1b);
    t2 = mul2(qw, L2E);
    const f32x2 e = pk2(ex2_mufu(t1a), ex2_mufu(t1b));
    const f32x2 lut = pk2(census_lut_ref(lut_base, p1a, p2a), census_lut_ref(lut_base, p1b, p2b));
    ct = add2(sub2(pk2(1.0f, 1.0f), e), lut);             // (1 - e) + census, two roundings like the scalar form
}
// range weights of FOUR samples (two pairs): w = __expf-equivalent(t2) * gg per value.  The `t2 < -126` fix-up of __expf (~1 % of the
// samples) is ONE test and one rarely taken block for the four; the values are formed as scalars so that the rare path patches them
// in place (patching a packed value costs two register copies on the common path).
__device__ __forceinline__ void sample_weight4(f32x2 t2a, f32x2 t2b, float g0, float g1, float g2, float g3, f32x2& wa, f32x2& wb) {
    float t[4], w[4];
    const float g[4] = {g0, g1, g2, g3};
    upk2(t2a, t[0], t[1]);
    upk2(t2b, t[2], t[3]);
#pragma unroll
    for (int k = 0; k < 4; k++) w[k] = __fmul_rn(ex2_mufu(t[k]), g[k]);
    if (fminf(fminf(t[0], t[1]), fminf(t[2], t[3])) < -126.0f) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (t[k] < -126.0f) w[k] = __fmul_rn(ex2_tiny(t[k]), g[k]);
    }
    wa = pk2(w[0], w[1]);
    wb = pk2(w[2], w[3]);
}
__device__ __forceinline__ void sample_weight4(f32x2 t2a, f32x2 t2b, float gg, f32x2& wa, f32x2& wb) { sample_weight4(t2a, t2b, gg, gg, gg, gg, wa, wb); }
// the same for one pair
__device__ __forceinline__ f32x2 sample_weight2(f32x2 t2, float g0, float g1) {
    float ta, tb;
    upk2(t2, ta, tb);
    float wa = __fmul_rn(ex2_mufu(ta), g0), wb = __fmul_rn(ex2_mufu(tb), g1);
    if (fminf(ta, tb) < -126.0f) {
        if (ta < -126.0f) wa = __fmul_rn(ex2_tiny(ta), g0);
        if (tb < -126.0f) wb = __fmul_rn(ex2_tiny(tb), g1);
    }
    return pk2(wa, wb);
}
__device__ __forceinline__ float min2(f32x2 v) { float a, b; upk2(v, a, b); return fminf(a, b); }

// sample_eval + the per-sample fix-up + accumulation: the plain form of one sample
__device__ __forceinline__ void sample_term(const float4& p1, const PixPk& p1k, const float4& p2, const PixPk& c2k, float d1, float gg,
                                            const float* s_census, float& cost_sum, float& weight_sum) {
    float cost, t2;
    sample_eval(p1, p1k, p2, c2k, d1, s_census, cost, t2);
    const bool tiny = t2 < -126.0f;   // __expf fix-up (see exp_ref)
    if (tiny) t2 = __fmul_rn(t2, 0.5f);
    float e2 = ex2_mufu(t2);
    if (tiny) e2 = __fmul_rn(e2, e2);
    const float w = __fmul_rn(e2, gg);
    cost_sum = __fmaf_rn(cost, w, cost_sum);
    weight_sum = __fadd_rn(weight_sum, w);
}

// G consecutive samples of ONE accumulator pair: evaluated side by side with the bare ex2, one test for the rare fix-up, then added
// in sample order -- the same bits as G calls of sample_term.
template <int G>
__device__ __forceinline__ void sample_group(const float4 (&p1)[G], const float4 (&p2)[G], const PixPk& c1k, const PixPk& c2k, const float (&gg)[G],
                                             const float* s_census, float& cost_sum, float& weight_sum) {
    float ct[G], t2[G], w[G];
    float tmin = 0.f;
#pragma unroll
    for (int k = 0; k < G; k++) {
        const PixPk p1k = pack_pix(p1[k]);
        const float d1 = max3abs_diff(c1k, p1k);
        sample_eval(p1[k], p1k, p2[k], c2k, d1, s_census, ct[k], t2[k]);
        w[k] = __fmul_rn(ex2_mufu(t2[k]), gg[k]);
        tmin = fminf(tmin, t2[k]);
    }
    if (tmin < -126.0f) {
#pragma unroll
        for (int k = 0; k < G; k++)
            if (t2[k] < -126.0f) w[k] = __fmul_rn(ex2_tiny(t2[k]), gg[k]);
    }
#pragma unroll
    for (int k = 0; k < G; k++) {
        cost_sum = __fmaf_rn(ct[k], w[k], cost_sum);
        weight_sum = __fadd_rn(weight_sum, w[k]);
    }
}

// _d_compute_patch_dist (bao_pmflow_kernel.cu:255-301): 100 samples at stride 2 over a 19x19 patch.
// A/B point at the PADDED origin of the source / target packed planes (same pitch).  TRANSPOSED = false: row-major planes
// (pixel (x,y) at y*pitch + x); true: column-major copies (pixel (x,y) at x*pitch + y) used by the row propagation
// passes so that lanes walking adjacent ROWS read adjacent addresses.  Sample order is i (rows) outer, j inner in both.
// Texture clamp addressing is realised by the replicated PAD border: |x2+j| never leaves it because
// targets lie in [0,w]x[0,h] and |j| <= 9 < PAD.
template <int STRIDE, bool TRANSPOSED>
__device__ __forceinline__ float patch_cost(const float4* __restrict__ A, const float4* __restrict__ B, int pitch, int x1, int y1, int x2, int y2,
                                            const CostLut& lut, const float* s_census) {
    // A/B: PADDED origin of the planes (pixel (-PAD,-PAD)); all offsets below are non-negative 32-bit element indices, so every
    // load address is one IMAD.WIDE.U32 away from a register-resident base pointer.
    const unsigned sj = TRANSPOSED ? pitch : 1, si = TRANSPOSED ? 1 : pitch;
    const unsigned oa = (unsigned)(x1 + PAD) * sj + (unsigned)(y1 + PAD) * si;
    const unsigned ob = (unsigned)(x2 + PAD) * sj + (unsigned)(y2 + PAD) * si;
    const PixPk c1k = pack_pix(ldpix(A + oa));
    const PixPk c2k = pack_pix(ldpix(B + ob));
    float cost_sum = 0.f, weight_sum = 0.f;
    constexpr int NJ = (2 * PATCH_R) / STRIDE + 1;           // samples per patch row
    constexpr int G = NJ % 5 == 0 ? 5 : (NJ == 7 ? 7 : 1);   // 10 -> two groups of 5, 7 -> one group, 19 -> plain
#pragma unroll 1
    for (int i = -PATCH_R; i <= PATCH_R; i += STRIDE) {
        const int ai = i < 0 ? -i : i;
        const unsigned ra = oa + (unsigned)(i * (int)si), rb = ob + (unsigned)(i * (int)si);
        if (G > 1) {
#pragma unroll
            for (int j0 = 0; j0 < NJ; j0 += G) {
                float4 p1[G], p2[G];
                float gg[G];
#pragma unroll
                for (int k = 0; k < G; k++) {
                    const int j = -PATCH_R + (j0 + k) * STRIDE;
                    p1[k] = ldpix(A + (ra + (unsigned)(j * (int)sj)));
                    p2[k] = ldpix(B + (rb + (unsigned)(j * (int)sj)));
                    gg[k] = lut.gg[ai][j < 0 ? -j : j];
                }
                sample_group<G>(p1, p2, c1k, c2k, gg, s_census, cost_sum, weight_sum);
            }
        } else {
#pragma unroll
            for (int j = -PATCH_R; j <= PATCH_R; j += STRIDE) {
                const float4 p1 = ldpix(A + (ra + (unsigned)(j * (int)sj)));
                const float4 p2 = ldpix(B + (rb + (unsigned)(j * (int)sj)));
                const PixPk p1k = pack_pix(p1);
                const float d1 = max3abs_diff(c1k, p1k);
                sample_term(p1, p1k, p2, c2k, d1, lut.gg[ai][j < 0 ? -j : j], s_census, cost_sum, weight_sum);
            }
        }
    }
    return __fdiv_rn(cost_sum, weight_sum);
}


// ---- parity-split ("Q") planes of the PatchMatch level --------------------------------------------------------------------------
// The patch samples every SECOND pixel in x and y (bao_pmflow_kernel.cu:269,272), so the 10 samples of a patch row are 32 bytes apart
// in a packed plane: one 16-byte request per sample and lane, half of every line fetched for nothing.  The Q plane stores the padded
// plane as four sub-planes by the parity of (x, y); the 100 samples of a patch then form a dense 10 x 10 block of ONE sub-plane and a
// lane reads two neighbouring samples with one 256-bit load (LDG.E.256, new on sm_100): half the L1 requests of the kernels that are
// bound by them.  A 256-bit load must be 32-byte aligned and a patch row starts at an arbitrary sample, so every sub-plane is stored
// twice, the second copy shifted by one sample (16 bytes): rows that start at an odd sample index read the shifted copy.
struct QGeom {
    int qp;          // row pitch of a sub-plane in pixels (even)
    int qh;          // rows of a sub-plane
    unsigned c1;     // element offset of the shifted copy (odd)
    unsigned plane;  // elements per image (both copies)
};
__host__ __device__ inline QGeom make_qgeom(int pw, int ph) {
    QGeom q;
    q.qp = ((pw + 1) / 2 + 2 + 1) & ~1;
    q.qh = (ph + 1) / 2 + 1;
    q.c1 = 4u * q.qh * q.qp + 3u;
    q.plane = (2u * 4u * q.qh * q.qp + 8u) & ~1u;
    return q;
}
// element index of padded pixel (X, Y) in copy 0
__device__ __forceinline__ unsigned q_index(const QGeom& q, int X, int Y) {
    return (unsigned)((((Y & 1) * 2 + (X & 1)) * q.qh + (Y >> 1)) * q.qp + (X >> 1));
}
struct Pix2 {
    float4 a, b;
};
__device__ __forceinline__ Pix2 ldpix2(const float4* p) {   // p 32-byte aligned
    Pix2 r;
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
        : "l"(p));
    return r;
}
// element index (even) of the top-left sample of the patch centred at logical pixel (x, y): padded (x + PAD - 9, y + PAD - 9)
__device__ __forceinline__ unsigned q_patch_origin(const QGeom& q, int x, int y) {
    const int X = x + PAD - PATCH_R, Y = y + PAD - PATCH_R;
    const unsigned e = q_index(q, X, Y);
    return (e & 1u) ? e + q.c1 : e;
}

// patch_cost at sample stride 2 on Q planes: QA / QB = Q planes of the source / target image (pair base), A / B the packed planes (centre
// pixels only).  Sample order, arithmetic and accumulation are patch_cost's: rows outer, samples inner.
__device__ __forceinline__ float patch_cost_q(const float4* __restrict__ A, const float4* __restrict__ B, const float4* __restrict__ QA,
                                              const float4* __restrict__ QB, const QGeom& q, int pw, int x1, int y1, int x2, int y2,
                                              const CostLut& lut, const float* s_census) {
    const PixPk c1k = pack_pix(ldpix(A + ((unsigned)(x1 + PAD) + (unsigned)(y1 + PAD) * (unsigned)pw)));
    const PixPk c2k = pack_pix(ldpix(B + ((unsigned)(x2 + PAD) + (unsigned)(y2 + PAD) * (unsigned)pw)));
    unsigned ea = q_patch_origin(q, x1, y1), eb = q_patch_origin(q, x2, y2);
    float cost_sum = 0.f, weight_sum = 0.f;
#pragma unroll 1
    for (int i = 0; i < 10; i++) {
        const int ai = i < 5 ? 9 - 2 * i : 2 * i - 9;
#pragma unroll
        for (int m = 0; m < 5; m++) {
            const Pix2 u = ldpix2(QA + (ea + 2u * m)), v = ldpix2(QB + (eb + 2u * m));
            const float4 p1[2] = {u.a, u.b}, p2[2] = {v.a, v.b};
            const int j0 = 2 * m < 5 ? 9 - 4 * m : 4 * m - 9, j1 = 2 * m + 1 < 5 ? 7 - 4 * m : 4 * m - 7;   // |-9 + 2 (2m)|, |-9 + 2 (2m + 1)|
            const float gg[2] = {lut.gg[ai][j0], lut.gg[ai][j1]};
            sample_group<2>(p1, p2, c1k, c2k, gg, s_census, cost_sum, weight_sum);
        }
        ea += q.qp;
        eb += q.qp;
    }
    return __fdiv_rn(cost_sum, weight_sum);
}

// mbarrier / TMA helpers (PTX, sm_90+): one thread arms the barrier with the byte count, issues the bulk tensor copy, everyone waits.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}


}  // namespace eppm
