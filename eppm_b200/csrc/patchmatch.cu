// Stage "patchmatch": randomised NNF at the coarsest pyramid level, both directions, whole batch per launch.
// Restates baoCudaPatchMatch (bao_pmflow_kernel.cu:1760-1826): baoGenerateRandomField (:50-109,153-165),
// baoComputeCostField (:636-645,689-696), baoSegPropagate (:1049-1181) and baoRandomSearch (:1519-1594).
//
// B200 design
//  * grid.z = pair*2 + direction: forward and backward fields of every pair of the batch advance in the same launch
//    (the reference runs 53 dependent launches per direction per pair on a 480x270 plane, i.e. ~5 warps per SM).
//  * RNG: the reference seeds XORWOW with 1234 and sub-sequence = 16x16-block id in both directions of every pair
//    (:68), so the random stream is a function of the level geometry alone.  It is expanded ONCE per context into
//    `rng_init` (initial targets) and `rng_search` (raw 16-bit draws per iteration/guess/pixel) by one thread per
//    sub-sequence; the per-pair kernels read it coalesced instead of serialising 3072 draws on one thread per block.
//  * propagation: all segments of a scan line live in one CTA and advance in lock-step with a CTA barrier per step,
//    which fixes the two orderings the reference leaves to warp scheduling (segment-start read before the
//    neighbour's last write; segment 1's first write to pixel 10 before segment 0's last) -- see DESIGN.md.
#include <curand_kernel.h>
#include <stdlib.h>

#include "eppm_internal.h"

namespace eppm {

constexpr int RB = 16;  // RNG block edge: BLOCK_DIM_X/Y of the reference (bao_pmflow_kernel.cu:42-43)

// One thread per 16x16 block id = per XORWOW sub-sequence.  Draw order follows d_gen_rand_field (:88-99: rows i, cols j,
// x then y) and d_update_random_guess (:1537-1551: per guess k, rows i, cols j, x then y), one search per iteration.
__global__ void k_rng_tables(short2* __restrict__ init, short2* __restrict__ search, int w, int h, int gx, int gy, int num_iter, int num_guess,
                             unsigned long long seed) {
    const int bid = blockIdx.x * blockDim.x + threadIdx.x;
    if (bid >= gx * gy) return;
    const int bx = bid % gx, by = bid / gx;
    curandState st;
    curand_init(seed, bid, 0, &st);  // :68
    for (int i = 0; i < RB; i++)
        for (int j = 0; j < RB; j++) {
            const unsigned r1 = curand(&st), r2 = curand(&st);
            const int x = bx * RB + j, y = by * RB + i;
            if (x < w && y < h) init[(size_t)y * w + x] = make_short2((short)(r1 % (unsigned)(w + 1)), (short)(r2 % (unsigned)(h + 1)));  // :97-98
        }
    for (int it = 0; it < num_iter; it++)
        for (int k = 0; k < num_guess; k++)
            for (int i = 0; i < RB; i++)
                for (int j = 0; j < RB; j++) {
                    const unsigned r1 = curand(&st), r2 = curand(&st);
                    const int x = bx * RB + j, y = by * RB + i;
                    if (x < w && y < h) search[((size_t)(it * num_guess + k) * h + y) * w + x] = make_short2((short)r1, (short)r2);  // :1549-1550
                }
}

// EPPM_RNG_PHILOX: counter-based Philox4x32-10 (curand's implementation of Salmon et al., SC'11), one independent stream per
// PIXEL: sub-sequence = y*w + x of the coarsest level, key = seed; draws 0,1 = initial target, then two draws per
// (iteration, guess).  Unlike XORWOW there is no serial dependence between pixels of a block; the same tables feed the same kernels.
__global__ void k_rng_tables_philox(short2* __restrict__ init, short2* __restrict__ search, int w, int h, int num_iter, int num_guess,
                                    unsigned long long seed) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= w * h) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)id, 0, &st);
    const unsigned r1 = curand(&st), r2 = curand(&st);
    init[id] = make_short2((short)(r1 % (unsigned)(w + 1)), (short)(r2 % (unsigned)(h + 1)));
    for (int it = 0; it < num_iter; it++)
        for (int k = 0; k < num_guess; k++) {
            const unsigned a = curand(&st), b = curand(&st);
            search[(size_t)(it * num_guess + k) * w * h + id] = make_short2((short)a, (short)b);
        }
}

void build_rng_tables(eppm_context* c) {
    const LevelGeom& g = c->lv[c->n_levels - 1];
    if (c->prm.rng_mode == EPPM_RNG_PHILOX) {
        k_rng_tables_philox<<<(g.w * g.h + 127) / 128, 128, 0, c->stream>>>(c->rng_init, c->rng_search, g.w, g.h, c->prm.num_iter, c->prm.num_rand_guess,
                                                                          c->prm.seed);
        EPPM_LAUNCH_COUNT(1);
        return;
    }
    const int gx = (g.w + RB - 1) / RB, gy = (g.h + RB - 1) / RB;
    k_rng_tables<<<(gx * gy + 63) / 64, 64, 0, c->stream>>>(c->rng_init, c->rng_search, g.w, g.h, gx, gy, c->prm.num_iter, c->prm.num_rand_guess,
                                                          c->prm.seed);
    EPPM_LAUNCH_COUNT(1);
}

void ensure_rng_tables(eppm_context* c) {
    if (c->rng_ready) return;
    if (!ensure_pm_buffers(c)) return;   // the launches that follow fail with the recorded error
    build_rng_tables(c);   // same stream as the PatchMatch kernels that follow
    c->rng_ready = 1;
}

// Residency cap of the PatchMatch kernels (co-scheduling with the refine kernel of another chunk, see eppm_compute_batch_host): a launch
// that asks for pad bytes of dynamic shared memory it never touches limits how many of its CTAs an SM holds, which leaves registers
// and warp slots to the kernel of the other stream.  0 = no cap.
static size_t pm_pad_bytes(eppm_context* c, const void* func) {
    const size_t pad = (size_t)c->pm_pad_kb * 1024;
    if (pad > 48 * 1024) cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
    return pad;
}

struct PmArgs {
    const float4* pix[2];   // packed planes of image 1 / image 2 at the PatchMatch level, padded origin of pair 0
    const float4* pixT[2];  // column-major copies (pixel (x,y) at (x+PAD)*ph + (y+PAD)), padded origin of pair 0
    unsigned plane;         // pixels per padded plane
    int pw, ph;
    short2* nnf[2];         // per direction, [B][h][w]
    float* cost[2];
    int w, h;
    int n_dirs;             // 2: grid.z = pair*2 + direction; 1: forward only (legacy single-direction entry point)
    int y0, y1;             // row band [y0, y1) this launch owns (whole level unless the frame is tiled across GPUs); segment aligned
    // optional texture path for scattered target-side gathers: linear uint4 textures over the packed planes of this level;
    // texB[dir] is the one that holds direction dir's TARGET image, texB_off[dir] the texel index of pair 0's padded origin in it
    const float4* q[2];     // parity-split (Q) planes of image 1 / image 2, pair 0 (null: not built)
    QGeom qg;
    int q_search;           // 1: the thread-per-pixel kernels (initial cost, random search, thread-mode queue scoring) read the Q planes too (EPPM_VAR_PM_Q)
    cudaTextureObject_t tex[2];      // per image: linear texture that holds its packed planes (0 = none)
    unsigned tex_off[2];             // texel index of pair 0's padded origin of pix[image] inside that texture
};

__device__ __forceinline__ float4 texpix(cudaTextureObject_t t, unsigned idx) {
    const uint4 v = tex1Dfetch<uint4>(t, (int)idx);
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}
template <bool T>
__device__ __forceinline__ void pm_select(const PmArgs& a, int z, const float4*& A, const float4*& B, short2*& nnf, float*& cost) {
    const int dir = a.n_dirs == 2 ? (z & 1) : 0, b = a.n_dirs == 2 ? (z >> 1) : z;
    // selects, not array indexing: a dynamically indexed kernel parameter is copied to local memory first
    const float4* i0 = T ? a.pixT[0] : a.pix[0];
    const float4* i1 = T ? a.pixT[1] : a.pix[1];
    A = (dir ? i1 : i0) + (size_t)b * a.plane;  // direction 1 swaps the images (…cuda.cpp:223-224)
    B = (dir ? i0 : i1) + (size_t)b * a.plane;
    nnf = (dir ? a.nnf[1] : a.nnf[0]) + (size_t)b * a.w * a.h;
    cost = (dir ? a.cost[1] : a.cost[0]) + (size_t)b * a.w * a.h;
    asm volatile("" : "+l"(A), "+l"(B));  // keep the plane bases in registers: every load is then base + u32 offset
}

__device__ __forceinline__ void pm_select_q(const PmArgs& a, int z, const float4*& QA, const float4*& QB) {
    const int dir = a.n_dirs == 2 ? (z & 1) : 0, b = a.n_dirs == 2 ? (z >> 1) : z;
    QA = (dir ? a.q[1] : a.q[0]) + (size_t)b * a.qg.plane;
    QB = (dir ? a.q[0] : a.q[1]) + (size_t)b * a.qg.plane;
    asm volatile("" : "+l"(QA), "+l"(QB));
}

// Random field + initial cost (d_gen_rand_field + d_compute_cost_field).
template <int STRIDE>
__global__ void __launch_bounds__(128) k_pm_init(PmArgs a, const short2* __restrict__ rng_init, const __grid_constant__ CostLut lut) {
    __shared__ float s_census[CENSUS_LUT_N];
    load_census_lut(s_census, lut);
    // CTA = blockDim.x consecutive pixels of blockDim.y consecutive rows (a 2-D tile: the patch windows of its pixels overlap in both directions)
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = a.y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= a.w || y >= a.y1) return;
    const float4 *A, *B; short2* nnf; float* cost;
    pm_select<false>(a, blockIdx.z, A, B, nnf, cost);
    const short2 t = rng_init[y * a.w + x];
    nnf[y * a.w + x] = t;
    if (STRIDE == 2 && a.q_search) {
        const float4 *QA, *QB;
        pm_select_q(a, blockIdx.z, QA, QB);
        cost[y * a.w + x] = patch_cost_q(A, B, QA, QB, a.qg, a.pw, x, y, t.x, t.y, lut, s_census);
    } else {
        cost[y * a.w + x] = patch_cost<STRIDE, false>(A, B, a.pw, x, y, t.x, t.y, lut, s_census);
    }
}

// Segment propagation, the four passes of baoSegPropagate.  DIR: 0 row forward, 1 column forward, 2 row reverse,
// 3 column reverse.  blockDim = (lines per CTA, all segments of a line); every thread owns one (line, segment).
// Lanes of a warp are ADJACENT scan lines working on the same position along the line: column passes read the
// row-major planes, row passes the column-major copies, so both sides of every sample are coalesced.
template <int DIR, int STRIDE>
__global__ void __launch_bounds__(896) k_pm_propagate(PmArgs a, int seg_len, int skip_equal, const __grid_constant__ CostLut lut) {
    constexpr bool ROW = (DIR == 0 || DIR == 2), FWD = (DIR < 2);
    __shared__ float s_census[CENSUS_LUT_N];
    load_census_lut(s_census, lut);
    // row passes: the band owns whole scan lines y0..y1-1; column passes: the band owns the segments that lie inside it
    const int line = (ROW ? a.y0 : 0) + blockIdx.x * blockDim.x + threadIdx.x;
    const int seg = (ROW ? 0 : a.y0 / seg_len) + threadIdx.y;
    const int n_line = ROW ? a.y1 : a.w;  // end of the scan lines
    const int len = ROW ? a.w : a.h;      // pixels along a line
    const float4 *A, *B; short2* nnf; float* cost;
    pm_select<ROW>(a, blockIdx.z, A, B, nnf, cost);
    const int pitch = ROW ? a.ph : a.pw;
    const bool active = line < n_line;
    int start, end, steps;
    if (FWD) {
        // :1055-1058  seg 0 starts at pixel 0 and covers one pixel more; others start one before their segment
        start = seg == 0 ? 0 : seg * seg_len - 1;
        end = min(len - 1, start + seg_len);
        steps = end - start;
    } else {
        // :1085-1088
        start = (seg + 1) * seg_len;
        if (start >= len) start = len - 1;
        end = seg * seg_len;
        steps = start - end;
    }
    if (!active) steps = 0;
    auto idx = [&](int i) -> int { return ROW ? line * a.w + i : i * a.w + line; };
    short2 prev = make_short2(0, 0);
    if (steps > 0) prev = nnf[idx(start)];
    __syncthreads();  // every segment has read its start pixel before any pixel is written
    for (int t = 1; t <= seg_len; t++) {
        if (t <= steps) {
            const int i = FWD ? start + t : start - t;
            const int id = idx(i);
            const float cur_best = cost[id];
            // :1065/:1095/:1125/:1155  shift the predecessor's target one step along the scan axis, clamped
            if (DIR == 0) prev.x = min(prev.x + 1, a.w - 1);
            if (DIR == 1) prev.y = min(prev.y + 1, a.h - 1);
            if (DIR == 2) prev.x = max(prev.x - 1, 0);
            if (DIR == 3) prev.y = max(prev.y - 1, 0);
            const int x1 = ROW ? i : line, y1 = ROW ? line : i;
            const short2 cur = nnf[id];
            // A candidate equal to the pixel's current target would be scored with the very evaluation that produced cost[id]
            // (same function, same arguments), so `cv < cur_best` is false: the reference's outcome without the 100 samples.
            if (skip_equal && prev.x == cur.x && prev.y == cur.y) {
                // prev stays (== nnf[id])
            } else {
                const float cv = patch_cost<STRIDE, ROW>(A, B, pitch, x1, y1, prev.x, prev.y, lut, s_census);
                if (cv < cur_best) {
                    nnf[id] = prev;
                    cost[id] = cv;
                } else {
                    prev = cur;
                }
            }
        }
        __syncthreads();  // lock-step: step t of every segment completes before step t+1 starts
    }
}

// The same passes with the evaluations COMPACTED inside the CTA.  From the third iteration on three quarters of the candidates equal
// the pixel's current target (measured at 1080p: 23 % of the threads need the 100 samples, yet 90 % of the warps contain one that
// does).  Per lock-step: every thread decides whether it needs an evaluation, the ones that do are packed into a queue in shared
// memory (ballot + per-warp counts), the first ceil(n/32) warps score the queue, owners pick up their result.  The propagation
// order, the candidates and the strict '<' are those of k_pm_propagate; only which lane computes a cost changes.
template <int DIR, int STRIDE>
__global__ void __launch_bounds__(896) k_pm_propagate_c(PmArgs a, int seg_len, const __grid_constant__ CostLut lut) {
    constexpr bool ROW = (DIR == 0 || DIR == 2), FWD = (DIR < 2);
    __shared__ float s_census[CENSUS_LUT_N];
    __shared__ int s_wcount[32];
    extern __shared__ int2 s_dyn[];
    const int nthreads = blockDim.x * blockDim.y;
    int2* s_item = s_dyn;                                        // queue: (owner thread | position along the line << 16, candidate target)
    float* s_res = reinterpret_cast<float*>(s_dyn + nthreads);   // cost of the owner's candidate
    load_census_lut(s_census, lut);
    const int tid = threadIdx.x + threadIdx.y * blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = (nthreads + 31) >> 5;
    const int line0 = (ROW ? a.y0 : 0) + blockIdx.x * blockDim.x;
    const int line = line0 + threadIdx.x;
    const int seg = (ROW ? 0 : a.y0 / seg_len) + threadIdx.y;
    const int n_line = ROW ? a.y1 : a.w;
    const int len = ROW ? a.w : a.h;
    const float4 *A, *B; short2* nnf; float* cost;
    pm_select<ROW>(a, blockIdx.z, A, B, nnf, cost);
    const int pitch = ROW ? a.ph : a.pw;
    const bool active = line < n_line;
    int start, end, steps;
    if (FWD) {
        start = seg == 0 ? 0 : seg * seg_len - 1;   // :1055-1058
        end = min(len - 1, start + seg_len);
        steps = end - start;
    } else {
        start = (seg + 1) * seg_len;                // :1085-1088
        if (start >= len) start = len - 1;
        end = seg * seg_len;
        steps = start - end;
    }
    if (!active) steps = 0;
    auto idx = [&](int l, int i) -> int { return ROW ? l * a.w + i : i * a.w + l; };
    short2 prev = make_short2(0, 0);
    if (steps > 0) prev = nnf[idx(line, start)];
    __syncthreads();  // every segment has read its start pixel before any pixel is written
    for (int t = 1; t <= seg_len; t++) {
        bool need = false;
        int i = 0, id = 0;
        short2 cur = make_short2(0, 0);
        if (t <= steps) {
            i = FWD ? start + t : start - t;
            id = idx(line, i);
            if (DIR == 0) prev.x = min(prev.x + 1, a.w - 1);   // :1065/:1095/:1125/:1155
            if (DIR == 1) prev.y = min(prev.y + 1, a.h - 1);
            if (DIR == 2) prev.x = max(prev.x - 1, 0);
            if (DIR == 3) prev.y = max(prev.y - 1, 0);
            cur = nnf[id];
            // a candidate equal to the current target would be scored by the evaluation that produced cost[id]: never '<'
            need = !(prev.x == cur.x && prev.y == cur.y);
        }
        const unsigned bal = __ballot_sync(__activemask(), need);
        if (lane == 0) s_wcount[warp] = __popc(bal);
        __syncthreads();
        int base = 0, total = 0;
        for (int w = 0; w < nwarps; w++) {
            const int cnt = s_wcount[w];
            if (w < warp) base += cnt;
            total += cnt;
        }
        if (need) s_item[base + __popc(bal & ((1u << lane) - 1))] = make_int2(tid | (i << 16), (int)(unsigned short)prev.x | ((int)prev.y << 16));
        __syncthreads();
        if (tid < total) {
            const int2 it = s_item[tid];
            const int owner = it.x & 0xffff, oi = it.x >> 16;
            const int ol = line0 + owner % (int)blockDim.x;
            const int x1 = ROW ? oi : ol, y1 = ROW ? ol : oi;
            s_res[owner] = patch_cost<STRIDE, ROW>(A, B, pitch, x1, y1, (short)(it.y & 0xffff), (short)(it.y >> 16), lut, s_census);
        }
        __syncthreads();
        if (need) {
            const float cv = s_res[tid];
            if (cv < cost[id]) {
                nnf[id] = prev;
                cost[id] = cv;
            } else {
                prev = cur;
            }
        }
        __syncthreads();  // lock-step: step t of every segment completes before step t+1 starts
    }
}

// ---- propagation as a global work queue (default) ----
// A lock-step of a pass is two launches over the WHOLE batch: k_prop_decide (one thread per (pair, direction, scan line, segment):
// shift the predecessor's target, compare with the pixel's own, enqueue the evaluations that are needed) and k_prop_eval (score
// the queue with every SM, apply the strict '<' update, leave the segment's running target in the per-thread state).  Inside one
// step all segments touch distinct pixels, and the launch boundary is the lock-step barrier, so the order of events is exactly
// k_pm_propagate's.  Measured at 1080p: from the third iteration on 77 % of the candidates equal the current target and are never
// scored; the CTA-local kernels cannot profit (their step time is the latency of one evaluation, however few lanes run it).
struct PropGeom {
    int n_line, line0;     // scan lines of the band and the first one
    int n_seg, seg0;       // segments per line handled here and the first one
    int len;               // pixels along a line
    int total;             // threads = n_z * n_seg * n_line
};

template <int DIR>
__global__ void __launch_bounds__(256) k_prop_decide(PmArgs a, PropGeom g, int seg_len, int t, short2* __restrict__ st_prev, int4* __restrict__ queue,
                                                     int* __restrict__ counter, int4* __restrict__ memo) {
    constexpr bool ROW = (DIR == 0 || DIR == 2), FWD = (DIR < 2);
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    bool need = false;
    int4 item = make_int4(0, 0, 0, 0);
    if (gid < g.total) {
        const int ll = gid % g.n_line, r = gid / g.n_line;
        const int seg = g.seg0 + r % g.n_seg, z = r / g.n_seg;
        const int line = g.line0 + ll;
        int start, steps;
        if (FWD) {
            start = seg == 0 ? 0 : seg * seg_len - 1;   // :1055-1058
            steps = min(g.len - 1, start + seg_len) - start;
        } else {
            start = (seg + 1) * seg_len;                // :1085-1088
            if (start >= g.len) start = g.len - 1;
            steps = start - seg * seg_len;
        }
        if (t <= steps) {
            const int dir = a.n_dirs == 2 ? (z & 1) : 0, b = a.n_dirs == 2 ? (z >> 1) : z;
            const short2* nnf = (dir ? a.nnf[1] : a.nnf[0]) + (size_t)b * a.w * a.h;
            short2 prev = t == 1 ? nnf[ROW ? line * a.w + start : start * a.w + line] : st_prev[gid];
            const int i = FWD ? start + t : start - t;
            if (DIR == 0) prev.x = min(prev.x + 1, a.w - 1);   // :1065/:1095/:1125/:1155
            if (DIR == 1) prev.y = min(prev.y + 1, a.h - 1);
            if (DIR == 2) prev.x = max(prev.x - 1, 0);
            if (DIR == 3) prev.y = max(prev.y - 1, 0);
            const int x1 = ROW ? i : line, y1 = ROW ? line : i;
            const short2 cur = nnf[y1 * a.w + x1];
            st_prev[gid] = prev;   // a skipped candidate equals the current target; an evaluated one is settled by k_prop_eval
            // a candidate equal to the current target would be scored by the evaluation that produced cost[id]: never '<'
            need = !(prev.x == cur.x && prev.y == cur.y);
            const int cand = (int)(unsigned short)prev.x | ((int)prev.y << 16);
            if (need && memo) {
                // A candidate this pixel has ALREADY scored can never win again: the cost of a (pixel, target) pair is a pure function of the
                // images, and the pixel's own cost only ever decreases, so `cv < cost` was false or made cost == cv then and is false now.
                // The memo keeps the last candidate scored at this pixel per pass direction (the same neighbour keeps proposing the same
                // target once the field has settled); a hit is the reference's outcome -- rejected, chain continues with the pixel's own
                // target -- without the 100 samples.
                int4* mp = memo + ((size_t)z * a.w * a.h + (size_t)y1 * a.w + x1);
                const int4 m = *mp;
                if (m.x == cand || m.y == cand || m.z == cand || m.w == cand) {
                    need = false;
                    st_prev[gid] = cur;
                } else {
                    reinterpret_cast<int*>(mp)[DIR] = cand;
                }
            }
            item = make_int4(z, x1 | (y1 << 16), cand, gid);
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, need);
    if (bal) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(counter, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (need) queue[base + __popc(bal & ((1u << lane) - 1))] = item;
    }
}

template <int DIR, int STRIDE, bool USEQ>
__global__ void __launch_bounds__(128) k_prop_eval(PmArgs a, const int4* __restrict__ queue, const int* __restrict__ counter, short2* __restrict__ st_prev,
                                                   const __grid_constant__ CostLut lut) {
    constexpr bool ROW = (DIR == 0 || DIR == 2);
    __shared__ float s_census[CENSUS_LUT_N];
    load_census_lut(s_census, lut);
    const int n = *counter;
    const int pitch = ROW ? a.ph : a.pw;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int4 it = queue[k];
        const float4 *A, *B; short2* nnf; float* cost;
        pm_select<ROW>(a, it.x, A, B, nnf, cost);
        const int x1 = it.y & 0xffff, y1 = it.y >> 16;
        const short2 cand = make_short2((short)(it.z & 0xffff), (short)(it.z >> 16));
        float cv;
        if (STRIDE == 2 && USEQ) {
            const float4 *A0, *B0, *QA, *QB; short2* n0; float* c0;
            pm_select<false>(a, it.x, A0, B0, n0, c0);
            pm_select_q(a, it.x, QA, QB);
            cv = patch_cost_q(A0, B0, QA, QB, a.qg, a.pw, x1, y1, cand.x, cand.y, lut, s_census);
        } else {
            cv = patch_cost<STRIDE, ROW>(A, B, pitch, x1, y1, cand.x, cand.y, lut, s_census);
        }
        const int id = y1 * a.w + x1;
        if (cv < cost[id]) {
            nnf[id] = cand;
            cost[id] = cv;
        } else {
            st_prev[it.w] = nnf[id];
        }
    }
}

// ---- warp-cooperative evaluation (default of the queue's scoring kernel and of the random search) ----
// A thread-per-evaluation kernel gathers 32 unrelated 19x19 windows per load: every lane hits a different cache line on both image
// sides (k_prop_eval: L1 hit rate 12 %, 53 % of the LSU wavefront peak at 27 % resident warps; the search spends one L1 wavefront per
// lane per sample).  Here the WARP scores one evaluation at a time: lane l takes samples l, l+32, l+64, ... of the patch, so a load
// covers ~3 sample rows of ONE window (10 samples of a row lie within 304 bytes: 3 lines instead of 10) and neighbouring evaluations
// share lines through L1.  Every sample's (cost, weight) goes to shared memory; after BATCH evaluations lane k adds up evaluation k
// IN SAMPLE ORDER (i outer, j inner, cs = fma(cost, w, cs), ws += w: bao_pmflow_kernel.cu:274-296), which is the reference's
// accumulation order and therefore its bits.  The serial sum costs 3 instructions per sample for BATCH evaluations at once.
template <int STRIDE>
struct Coop {
    static constexpr int NJ = (2 * PATCH_R) / STRIDE + 1;
    static constexpr int NS = NJ * NJ;            // samples per evaluation (100 at stride 2)
    static constexpr int NSP = NS | 1;            // odd pitch (in float2) of an evaluation in shared memory: the 8-byte reads of the serial sums fall in distinct banks
    static constexpr int R = (NS + 31) / 32;      // rounds of 32 samples
};

// per-lane sample sites: sample s = lane + 32 r lies at (i, j) = (-9 + STRIDE * (s / NJ), -9 + STRIDE * (s % NJ))
template <int STRIDE>
__device__ __forceinline__ void coop_sites(int lane, int pw, const CostLut& lut, int (&soff)[Coop<STRIDE>::R], float (&sgg)[Coop<STRIDE>::R]) {
    typedef Coop<STRIDE> C;
#pragma unroll
    for (int r = 0; r < C::R; r++) {
        int s = lane + 32 * r;
        if (s >= C::NS) s = C::NS - 1;   // lanes past the last sample score a valid site and do not store
        const int i = -PATCH_R + STRIDE * (s / C::NJ), j = -PATCH_R + STRIDE * (s % C::NJ);
        soff[r] = i * pw + j;
        sgg[r] = lut.gg[i < 0 ? -i : i][j < 0 ? -j : j];
    }
}

// image-1 side of an evaluation, per lane: the R samples of the source patch and their range distances to its centre
template <int STRIDE>
struct CoopSrc {
    float4 p1[Coop<STRIDE>::R];
    float d1[Coop<STRIDE>::R];
};
template <int STRIDE>
__device__ __forceinline__ void coop_load_src(CoopSrc<STRIDE>& S, const float4* __restrict__ A, unsigned oa, const int (&soff)[Coop<STRIDE>::R]) {
    const PixPk c1k = pack_pix(ldpix(A + oa));
#pragma unroll
    for (int r = 0; r < Coop<STRIDE>::R; r++) S.p1[r] = ldpix(A + (oa + (unsigned)soff[r]));
#pragma unroll
    for (int r = 0; r < Coop<STRIDE>::R; r++) S.d1[r] = max3abs_diff(c1k, pack_pix(S.p1[r]));
}
// the lane's samples of ONE evaluation against target offset ob, written to `slot`
template <int STRIDE>
__device__ __forceinline__ void coop_score(const CoopSrc<STRIDE>& S, const float4* __restrict__ B, unsigned ob, const int (&soff)[Coop<STRIDE>::R],
                                           const float (&sgg)[Coop<STRIDE>::R], unsigned lut_base, int lane, float2* __restrict__ slot) {
    typedef Coop<STRIDE> C;
    const PixPk c2k = pack_pix(ldpix(B + ob));
    float4 p2[C::R];
#pragma unroll
    for (int r = 0; r < C::R; r++) p2[r] = ldpix(B + (ob + (unsigned)soff[r]));
    float ct[C::R], t2[C::R], w[C::R];
    float tmin = 0.f;
#pragma unroll
    for (int r = 0; r < C::R; r++) {
        sample_eval(S.p1[r], pack_pix(S.p1[r]), p2[r], c2k, S.d1[r], lut_base, ct[r], t2[r]);
        w[r] = __fmul_rn(ex2_mufu(t2[r]), sgg[r]);
        tmin = fminf(tmin, t2[r]);
    }
    if (tmin < -126.0f) {   // the rare __expf fix-up, one test for the lane's samples (see sample_group)
#pragma unroll
        for (int r = 0; r < C::R; r++)
            if (t2[r] < -126.0f) w[r] = __fmul_rn(ex2_tiny(t2[r]), sgg[r]);
    }
#pragma unroll
    for (int r = 0; r < C::R; r++)
        if ((r + 1) * 32 <= C::NS || lane + 32 * r < C::NS) slot[lane + 32 * r] = make_float2(ct[r], w[r]);
}
// the serial sum of one evaluation (the owner lane): the reference's accumulation order
template <int STRIDE>
__device__ __forceinline__ float coop_sum(const float2* __restrict__ slot) {
    float cs = 0.f, ws = 0.f;
#pragma unroll 10
    for (int s = 0; s < Coop<STRIDE>::NS; s++) {
        const float2 v = slot[s];
        cs = __fmaf_rn(v.x, v.y, cs);
        ws = __fadd_rn(ws, v.y);
    }
    return __fdiv_rn(cs, ws);
}

// k_prop_eval with warp-cooperative scoring: a warp takes BATCH consecutive queue items
template <int STRIDE, int BATCH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_prop_eval_w(PmArgs a, const int4* __restrict__ queue, const int* __restrict__ counter, short2* __restrict__ st_prev,
                                                           const __grid_constant__ CostLut lut) {
    typedef Coop<STRIDE> C;
    __shared__ float s_census[CENSUS_LUT_N];
    extern __shared__ float2 s_val_dyn[];   // [WARPS][BATCH][NSP]
    load_census_lut(s_census, lut);
    const unsigned lut_base = census_lut_base(s_census);
    const int n = *counter;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* s_val = s_val_dyn + (size_t)warp * BATCH * C::NSP;
    int soff[C::R];
    float sgg[C::R];
    coop_sites<STRIDE>(lane, a.pw, lut, soff, sgg);
    for (int base = (blockIdx.x * WARPS + warp) * BATCH; base < n; base += gridDim.x * WARPS * BATCH) {
        const int cnt = min(BATCH, n - base);
        int4 it = make_int4(0, 0, 0, 0);
        if (lane < cnt) it = queue[base + lane];
        for (int k = 0; k < cnt; k++) {
            const int z = __shfl_sync(0xffffffffu, it.x, k), pos = __shfl_sync(0xffffffffu, it.y, k), cand = __shfl_sync(0xffffffffu, it.z, k);
            const float4 *A, *B; short2* nnf; float* cost;
            pm_select<false>(a, z, A, B, nnf, cost);
            const unsigned oa = (unsigned)((pos & 0xffff) + PAD) + (unsigned)((pos >> 16) + PAD) * (unsigned)a.pw;
            const unsigned ob = (unsigned)((short)(cand & 0xffff) + PAD) + (unsigned)((cand >> 16) + PAD) * (unsigned)a.pw;
            CoopSrc<STRIDE> S;
            coop_load_src<STRIDE>(S, A, oa, soff);
            coop_score<STRIDE>(S, B, ob, soff, sgg, lut_base, lane, s_val + k * C::NSP);
        }
        __syncwarp();
        if (lane < cnt) {
            const float cv = coop_sum<STRIDE>(s_val + lane * C::NSP);
            const float4 *A, *B; short2* nnf; float* cost;
            pm_select<false>(a, it.x, A, B, nnf, cost);
            const int id = (it.y >> 16) * a.w + (it.y & 0xffff);
            if (cv < cost[id]) {
                nnf[id] = make_short2((short)(it.z & 0xffff), (short)(it.z >> 16));
                cost[id] = cv;
            } else {
                st_prev[it.w] = nnf[id];
            }
        }
        __syncwarp();
    }
}

// Round-major form of k_prop_eval_w (default).  Staging all samples of BATCH evaluations takes 13 KB of shared memory per warp; at
// four warps per CTA the occupancy limit fills the whole 228 KB carve-out, leaves the L1 cache ~30 KB and the kernel waits on L2
// (measured slower than the thread-per-evaluation kernel).  Here a warp walks its BATCH items once per ROUND of 32 samples: stage
// (cost, weight) of that round only (BATCH x 33 float2 = 4 KB per warp), the owner lanes fold the 32 samples into their running
// (cost_sum, weight_sum) registers -- still sample order -- and the next round reuses the buffer.  The loads of U items are issued
// together (4 x U 16-byte loads in flight per lane).  The samples past the last full round (4 of 100 at stride 2) are scored for
// 32 / 4 items at once, lane l = (item l / 4, sample l % 4), instead of a round with 28 idle lanes per item.
template <int STRIDE, int BATCH, int WARPS, bool USEQ>
__global__ void __launch_bounds__(WARPS * 32) k_prop_eval_r(PmArgs a, const int4* __restrict__ queue, const int* __restrict__ counter, short2* __restrict__ st_prev,
                                                           const __grid_constant__ CostLut lut) {
    typedef Coop<STRIDE> C;
    constexpr int RF = C::NS / 32;             // full rounds
    constexpr int TS = C::NS - 32 * RF;        // samples of the tail
    constexpr int IPP = TS > 0 ? 32 / TS : 1;  // items per tail pass
    constexpr int U = 4;
    constexpr int PITCH = 33;
    __shared__ float s_census[CENSUS_LUT_N];
    __shared__ float2 s_stage[WARPS][BATCH][PITCH];
    load_census_lut(s_census, lut);
    const unsigned lut_base = census_lut_base(s_census);
    const int n = *counter;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 (*stage)[PITCH] = s_stage[warp];
    // USEQ: the samples are read from the parity-split (Q) planes, where the 10 x 10 samples of a patch are a dense block of ONE sub-plane
    // (sample (i, j) at origin + i * qp + j): a warp-wide request covers 3.2 sample rows of 160 contiguous bytes instead of 304-byte
    // spans with every second pixel unused -- half the L1 wavefronts per request.  Centre pixels still come from the packed planes.
    int soff[C::R];
    float sgg[C::R];
    coop_sites<STRIDE>(lane, a.pw, lut, soff, sgg);
    if (USEQ) {
#pragma unroll
        for (int r = 0; r < C::R; r++) {
            const int sx = min(lane + 32 * r, C::NS - 1);
            soff[r] = (sx / C::NJ) * a.qg.qp + (sx % C::NJ);   // relative to the Q element of the patch's top-left sample
        }
    }
    // tail site of this lane: sample 32 RF + lane % TS
    int toff = 0;
    float tgg = 0.f;
    if (TS > 0) {
        const int s = 32 * RF + lane % (TS > 0 ? TS : 1);
        const int i = -PATCH_R + STRIDE * (s / C::NJ), j = -PATCH_R + STRIDE * (s % C::NJ);
        toff = USEQ ? (s / C::NJ) * a.qg.qp + (s % C::NJ) : i * a.pw + j;
        tgg = lut.gg[i < 0 ? -i : i][j < 0 ? -j : j];
    }
    for (int base = (blockIdx.x * WARPS + warp) * BATCH; base < n; base += gridDim.x * WARPS * BATCH) {
        const int cnt = min(BATCH, n - base);
        int4 it = make_int4(0, 0, 0, 0);
        if (lane < cnt) it = queue[base + lane];
        float cs = 0.f, ws = 0.f;
#pragma unroll
        for (int r = 0; r < RF; r++) {
            for (int k0 = 0; k0 < cnt; k0 += U) {
                float4 c1[U], c2[U], p1[U], p2[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int k = min(k0 + u, cnt - 1);   // past the end: the last item again (same values to the same slot)
                    const int z = __shfl_sync(0xffffffffu, it.x, k), pos = __shfl_sync(0xffffffffu, it.y, k), cand = __shfl_sync(0xffffffffu, it.z, k);
                    const float4 *A, *B; short2* nnf; float* cost;
                    pm_select<false>(a, z, A, B, nnf, cost);
                    const unsigned oa = (unsigned)((pos & 0xffff) + PAD) + (unsigned)((pos >> 16) + PAD) * (unsigned)a.pw;
                    const unsigned ob = (unsigned)((short)(cand & 0xffff) + PAD) + (unsigned)((cand >> 16) + PAD) * (unsigned)a.pw;
                    c1[u] = ldpix(A + oa);
                    c2[u] = ldpix(B + ob);
                    if (USEQ) {
                        const float4 *QA, *QB;
                        pm_select_q(a, z, QA, QB);
                        p1[u] = ldpix(QA + (q_index(a.qg, (pos & 0xffff) + PAD - PATCH_R, (pos >> 16) + PAD - PATCH_R) + (unsigned)soff[r]));
                        p2[u] = ldpix(QB + (q_index(a.qg, (short)(cand & 0xffff) + PAD - PATCH_R, (cand >> 16) + PAD - PATCH_R) + (unsigned)soff[r]));
                    } else {
                        p1[u] = ldpix(A + (oa + (unsigned)soff[r]));
                        p2[u] = ldpix(B + (ob + (unsigned)soff[r]));
                    }
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int k = min(k0 + u, cnt - 1);
                    const PixPk p1k = pack_pix(p1[u]);
                    float ct, t2;
                    sample_eval(p1[u], p1k, p2[u], pack_pix(c2[u]), max3abs_diff(pack_pix(c1[u]), p1k), lut_base, ct, t2);
                    float w = __fmul_rn(ex2_mufu(t2), sgg[r]);
                    if (t2 < -126.0f) w = __fmul_rn(ex2_tiny(t2), sgg[r]);
                    stage[k][lane] = make_float2(ct, w);
                }
            }
            __syncwarp();
            if (lane < cnt) {
#pragma unroll 8
                for (int s = 0; s < 32; s++) {
                    const float2 v = stage[lane][s];
                    cs = __fmaf_rn(v.x, v.y, cs);
                    ws = __fadd_rn(ws, v.y);
                }
            }
            __syncwarp();
        }
        if (TS > 0) {
            for (int g0 = 0; g0 < cnt; g0 += IPP) {
                const int ki = g0 + lane / TS;
                const bool valid = lane < IPP * TS && ki < cnt;
                const int src = min(ki, cnt - 1);
                const int z = __shfl_sync(0xffffffffu, it.x, src), pos = __shfl_sync(0xffffffffu, it.y, src), cand = __shfl_sync(0xffffffffu, it.z, src);
                const float4 *A, *B; short2* nnf; float* cost;
                pm_select<false>(a, z, A, B, nnf, cost);
                const unsigned oa = (unsigned)((pos & 0xffff) + PAD) + (unsigned)((pos >> 16) + PAD) * (unsigned)a.pw;
                const unsigned ob = (unsigned)((short)(cand & 0xffff) + PAD) + (unsigned)((cand >> 16) + PAD) * (unsigned)a.pw;
                const float4 c1 = ldpix(A + oa), c2 = ldpix(B + ob);
                float4 p1, p2;
                if (USEQ) {
                    const float4 *QA, *QB;
                    pm_select_q(a, z, QA, QB);
                    p1 = ldpix(QA + (q_index(a.qg, (pos & 0xffff) + PAD - PATCH_R, (pos >> 16) + PAD - PATCH_R) + (unsigned)toff));
                    p2 = ldpix(QB + (q_index(a.qg, (short)(cand & 0xffff) + PAD - PATCH_R, (cand >> 16) + PAD - PATCH_R) + (unsigned)toff));
                } else {
                    p1 = ldpix(A + (oa + (unsigned)toff));
                    p2 = ldpix(B + (ob + (unsigned)toff));
                }
                const PixPk p1k = pack_pix(p1);
                float ct, t2;
                sample_eval(p1, p1k, p2, pack_pix(c2), max3abs_diff(pack_pix(c1), p1k), lut_base, ct, t2);
                float w = __fmul_rn(ex2_mufu(t2), tgg);
                if (t2 < -126.0f) w = __fmul_rn(ex2_tiny(t2), tgg);
                if (valid) stage[ki][lane % TS] = make_float2(ct, w);
            }
            __syncwarp();
            if (lane < cnt) {
#pragma unroll
                for (int s = 0; s < TS; s++) {
                    const float2 v = stage[lane][s];
                    cs = __fmaf_rn(v.x, v.y, cs);
                    ws = __fadd_rn(ws, v.y);
                }
            }
        }
        if (lane < cnt) {
            const float cv = __fdiv_rn(cs, ws);
            const float4 *A, *B; short2* nnf; float* cost;
            pm_select<false>(a, it.x, A, B, nnf, cost);
            const int id = (it.y >> 16) * a.w + (it.y & 0xffff);
            if (cv < cost[id]) {
                nnf[id] = make_short2((short)(it.z & 0xffff), (short)(it.z >> 16));
                cost[id] = cv;
            } else {
                st_prev[it.w] = nnf[id];
            }
        }
        __syncwarp();
    }
}

// Random search, warp-cooperative: a warp owns PIX consecutive pixels of a row and scores their PIX * NG guesses one after the
// other (the image-1 side of a pixel's samples is loaded once for its NG guesses); lane q = (pixel q / NG, guess q % NG) then adds
// up evaluation q in sample order, and the guesses of a pixel are compared in their order with strict '<' (:1577).
template <int STRIDE, int NG, int PIX, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_pm_search_w(PmArgs a, const short2* __restrict__ rng, int search_range, int radius_min,
                                                           const __grid_constant__ CostLut lut) {
    typedef Coop<STRIDE> C;
    static_assert(PIX * NG <= 32, "one lane per (pixel, guess)");
    __shared__ float s_census[CENSUS_LUT_N];
    extern __shared__ float2 s_val_dyn[];   // [WARPS][PIX * NG][NSP]
    load_census_lut(s_census, lut);
    const unsigned lut_base = census_lut_base(s_census);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* s_val = s_val_dyn + (size_t)warp * (PIX * NG) * C::NSP;
    const int x0 = (blockIdx.x * WARPS + warp) * PIX, y = a.y0 + blockIdx.y;
    if (x0 >= a.w) return;   // whole warps only; nothing below synchronises across warps
    const float4 *A, *B; short2* nnf; float* cost;
    pm_select<false>(a, blockIdx.z, A, B, nnf, cost);
    int soff[C::R];
    float sgg[C::R];
    coop_sites<STRIDE>(lane, a.pw, lut, soff, sgg);
    // lane q: its pixel, its guess
    const int qp = lane / NG, qk = lane - qp * NG;
    const int x = x0 + qp;
    const bool mine = lane < PIX * NG && x < a.w;
    const int id = y * a.w + min(x, a.w - 1);
    const short2 entry = nnf[id];
    int cand = 0;
    {
        int mag = search_range;
        for (int k = 0; k < qk; k++)
            if (mag / 2 >= radius_min) mag /= 2;
        const short2 rr = rng[(size_t)qk * a.w * a.h + id];
        const unsigned r1 = (unsigned)(int)rr.x, r2 = (unsigned)(int)rr.y;   // :1557-1563
        const short xmin = (short)max(entry.x - mag, 0), xmax = (short)min(entry.x + mag + 1, a.w + 1);
        const short ymin = (short)max(entry.y - mag, 0), ymax = (short)min(entry.y + mag + 1, a.h + 1);
        const short gx = (short)(xmin + r1 % (unsigned)(xmax - xmin));
        const short gy = (short)(ymin + r2 % (unsigned)(ymax - ymin));
        cand = (int)(unsigned short)gx | ((int)gy << 16);
    }
    const int npix = min(PIX, a.w - x0);
    CoopSrc<STRIDE> S;
    for (int q = 0; q < npix * NG; q++) {
        if (q % NG == 0) {
            const unsigned oa = (unsigned)(x0 + q / NG + PAD) + (unsigned)(y + PAD) * (unsigned)a.pw;
            coop_load_src<STRIDE>(S, A, oa, soff);
        }
        const int ec = __shfl_sync(0xffffffffu, cand, q);
        const unsigned ob = (unsigned)((short)(ec & 0xffff) + PAD) + (unsigned)((ec >> 16) + PAD) * (unsigned)a.pw;
        coop_score<STRIDE>(S, B, ob, soff, sgg, lut_base, lane, s_val + q * C::NSP);
    }
    __syncwarp();
    float cv = 0.f;
    if (mine) cv = coop_sum<STRIDE>(s_val + lane * C::NSP);
    // the guesses of a pixel in their order (every lane of the pixel follows the scan; the lane of guess 0 stores)
    int best = (int)(unsigned short)entry.x | ((int)entry.y << 16);
    float best_cost = cost[id];
#pragma unroll
    for (int k = 0; k < NG; k++) {
        const int src = min(qp * NG + k, 31);
        const float cvk = __shfl_sync(0xffffffffu, cv, src);
        const int ck = __shfl_sync(0xffffffffu, cand, src);
        if (cvk < best_cost) {
            best = ck;
            best_cost = cvk;
        }
    }
    if (mine && qk == 0) {
        nnf[id] = make_short2((short)(best & 0xffff), (short)(best >> 16));
        cost[id] = best_cost;
    }
}

template <int DIR, int STRIDE>
static void launch_propagate_queue(eppm_context* c, const PmArgs& a, int n, int pass_index) {
    const bool row = (DIR == 0 || DIR == 2);
    const int sl = c->prm.prop_seg_length;
    PropGeom g;
    g.n_line = row ? a.y1 - a.y0 : a.w;
    g.line0 = row ? a.y0 : 0;
    g.n_seg = row ? (a.w + sl - 1) / sl : (a.y1 + sl - 1) / sl - a.y0 / sl;
    g.seg0 = row ? 0 : a.y0 / sl;
    g.len = row ? a.w : a.h;
    g.total = a.n_dirs * n * g.n_seg * g.n_line;
    int* counters = c->prop_count + (size_t)pass_index * sl;
    cudaMemsetAsync(counters, 0, sizeof(int) * sl, c->stream);
    // evaluation grid: one thread per possible item by default (CTAs beyond the queue length exit at once; the hardware hands CTAs to SMs as
    // they drain, which balances better than a capped grid striding over the queue: 5.06 -> 5.01 ms per 1080p pair); EPPM_PROP_EVAL_CAP = CTAs per SM
    static const int cap = getenv("EPPM_PROP_EVAL_CAP") ? atoi(getenv("EPPM_PROP_EVAL_CAP")) : 0;
    const int eval_blocks = cap > 0 ? min((g.total + 127) / 128, c->n_sm * cap) : (g.total + 127) / 128;
    // warp-cooperative scoring: round-major (default) or whole evaluations staged (EPPM_VAR_PROP_WARP_FULL)
    constexpr int WB = STRIDE == 1 ? 4 : 16, WW = 4;
    constexpr int RW = 4;
    const size_t wsmem = (size_t)WW * WB * Coop<STRIDE>::NSP * sizeof(float2);
    static bool attr_w[64] = {};
    const int mode = (c->variant & EPPM_VAR_PROP_THREAD) ? 0 : (c->variant & EPPM_VAR_PROP_WARP_FULL) ? 1 : 2;
    static const int env_batch = getenv("EPPM_PROP_BATCH") ? atoi(getenv("EPPM_PROP_BATCH")) : 8;   // tuning knob: queue items per warp (8: 4.20, 16: 4.24 ms of PatchMatch per pair)
    const bool rq = STRIDE == 2 && a.q[0] != nullptr;   // Q planes built: the round-major kernel reads its samples from them
    if (mode && (c->device < 0 || c->device >= 64 || !attr_w[c->device])) {
        cudaFuncSetAttribute(k_prop_eval_w<STRIDE, WB, WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
        // leave the larger part of the unified L1 / shared memory to the cache: the patch windows of neighbouring queue items overlap
        static const int carve = getenv("EPPM_PROP_CARVEOUT") ? atoi(getenv("EPPM_PROP_CARVEOUT")) : 40;
        cudaFuncSetAttribute(k_prop_eval_r<STRIDE, 16, RW, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_prop_eval_r<STRIDE, 8, RW, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_prop_eval_r<STRIDE, 16, RW, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_prop_eval_r<STRIDE, 8, RW, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (c->device >= 0 && c->device < 64) attr_w[c->device] = true;
    }
    const int rb = env_batch == 16 ? 16 : 8;
    const int wblocks = (g.total + WB * WW - 1) / (WB * WW), rblocks = (g.total + rb * RW - 1) / (rb * RW);
    for (int t = 1; t <= sl; t++) {
        k_prop_decide<DIR><<<(g.total + 255) / 256, 256, 0, c->stream>>>(a, g, sl, t, c->prop_prev, c->prop_queue, counters + t - 1,
                                                                         (c->variant & EPPM_VAR_PROP_NOMEMO) ? nullptr : c->prop_memo);
#define EPPM_EVR(B, Q) k_prop_eval_r<STRIDE, B, RW, Q><<<rblocks, RW * 32, 0, c->stream>>>(a, c->prop_queue, counters + t - 1, c->prop_prev, c->cost_lut)
        if (mode == 2) {
            if (rb == 8) { if (rq) EPPM_EVR(8, true); else EPPM_EVR(8, false); }
            else { if (rq) EPPM_EVR(16, true); else EPPM_EVR(16, false); }
        }
#undef EPPM_EVR
        else if (mode == 1) k_prop_eval_w<STRIDE, WB, WW><<<wblocks, WW * 32, wsmem, c->stream>>>(a, c->prop_queue, counters + t - 1, c->prop_prev, c->cost_lut);
        else if (STRIDE == 2 && a.q_search) k_prop_eval<DIR, STRIDE, true><<<eval_blocks, 128, 0, c->stream>>>(a, c->prop_queue, counters + t - 1, c->prop_prev, c->cost_lut);
        else k_prop_eval<DIR, STRIDE, false><<<eval_blocks, 128, pm_pad_bytes(c, (const void*)k_prop_eval<DIR, STRIDE, false>), c->stream>>>(a, c->prop_queue, counters + t - 1, c->prop_prev, c->cost_lut);
    }
    EPPM_LAUNCH_COUNT(2 * sl);
}

// ---- propagation as independent segment chains with warp-shared evaluations (default) ----
// What the lock-step really orders.  A segment touches only its own pixels; the cost of a candidate reads the (immutable) images
// and the pixel's own cost.  Across segments of a scan line exactly three things are ordered by the reference's warp-synchronous
// execution (and by the barrier / launch boundary per step of the kernels above):
//   (1) every segment reads the pixel in front of it -- the last pixel of the neighbouring segment -- before anything is written;
//   (2) forward passes: segments 0 and 1 both update pixel seg_len (:1055-1058); segment 1 does it in its FIRST step, segment 0 in
//       its LAST one, which therefore sees segment 1's result;
//   (3) nothing else: reverse segments are disjoint.
// So a pass is: k_prop_snapshot (the start targets of all chains, ordering (1) by a launch boundary), ONE kernel in which every
// chain runs its steps back to back with no barrier at all, and for forward passes a tail launch with the last step of segment 0
// (ordering (2)).  3 launches per pass instead of 20, and no grid-wide wait per step.
//
// Inside the chain kernel a warp owns 32 chains (adjacent scan lines, same segment).  Per step the lanes shift their targets and
// decide (a candidate equal to the pixel's current target is never better, see k_pm_propagate); the evaluations that are needed are
// then done BY THE WHOLE WARP, one after the other: lane l scores samples l, l+32, l+64, l+96 of the patch (both sides of every
// sample are short row segments of a 19x19 window: coalesced, and shared through L1 with the neighbouring lines' evaluations; the
// thread-per-evaluation kernels gather 32 unrelated windows per load, L1 hit rate 12 %), writes (cost, weight) per sample to shared
// memory, and the owning lane adds the samples up IN SAMPLE ORDER -- the reference's accumulation order, hence its bits.  The four
// rounds of a lane are two packed pairs (sample_eval2).
struct ChainGeom {
    int n_line, line0, nlp;   // scan lines of the band, the first one, n_line rounded up to a multiple of 32 (a warp never straddles segments)
    int n_seg, seg0;          // segments per line handled by this context (band) and the first one
    int seg_count;            // segments enumerated by this launch: n_seg (main) or 1 (tail: segment 0 only)
    int len;                  // pixels along a line
    int t0, t1;               // steps [t0, t1] of every chain run in this launch
    int defer;                // 1: segment 0 stops one step early (its last step runs in the tail launch)
    int n_z;                  // pairs x directions
};

template <int DIR>
__device__ __forceinline__ void chain_span(int seg, int seg_len, int len, int& start, int& steps) {
    if (DIR < 2) {
        start = seg == 0 ? 0 : seg * seg_len - 1;   // :1055-1058
        steps = min(len - 1, start + seg_len) - start;
    } else {
        start = (seg + 1) * seg_len;                // :1085-1088
        if (start >= len) start = len - 1;
        steps = start - seg * seg_len;
    }
}

template <int DIR>
__global__ void __launch_bounds__(256) k_prop_snapshot(PmArgs a, ChainGeom g, int seg_len, short2* __restrict__ st_prev) {
    constexpr bool ROW = (DIR == 0 || DIR == 2);
    const int cid = blockIdx.x * blockDim.x + threadIdx.x;
    if (cid >= g.n_z * g.n_seg * g.nlp) return;
    const int ll = cid % g.nlp, r = cid / g.nlp;
    const int seg = g.seg0 + r % g.n_seg, z = r / g.n_seg;
    if (ll >= g.n_line) return;
    const int line = g.line0 + ll;
    int start, steps;
    chain_span<DIR>(seg, seg_len, g.len, start, steps);
    if (steps <= 0) return;
    const int dir = a.n_dirs == 2 ? (z & 1) : 0, b = a.n_dirs == 2 ? (z >> 1) : z;
    const short2* nnf = (dir ? a.nnf[1] : a.nnf[0]) + (size_t)b * a.w * a.h;
    st_prev[cid] = nnf[ROW ? line * a.w + start : start * a.w + line];
}

template <int STRIDE>
struct ChainCfg {
    static constexpr int NJ = (2 * PATCH_R) / STRIDE + 1;
    static constexpr int NS = NJ * NJ;                 // samples per evaluation
    static constexpr int NSP = NS | 1;                 // odd pitch of a slot in shared memory: the owners' 8-byte reads fall in distinct banks
    static constexpr int R = (NS + 31) / 32;           // rounds of 32 samples
    static constexpr int RP = (R + 1) / 2;             // rounds in packed pairs
    static constexpr int SLOTS = STRIDE == 1 ? 2 : 8;  // evaluations buffered per warp before their owners add them up
    static constexpr int WARPS = 4;
};

template <int DIR, int STRIDE>
__global__ void __launch_bounds__(ChainCfg<STRIDE>::WARPS * 32) k_prop_chain(PmArgs a, ChainGeom g, int seg_len, short2* __restrict__ st_prev,
                                                                            const __grid_constant__ CostLut lut) {
    typedef ChainCfg<STRIDE> C;
    constexpr bool ROW = (DIR == 0 || DIR == 2), FWD = (DIR < 2);
    __shared__ float s_census[CENSUS_LUT_N];
    __shared__ float2 s_val[C::WARPS][C::SLOTS][C::NSP];
    load_census_lut(s_census, lut);
    const unsigned lut_base = census_lut_base(s_census);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wi = blockIdx.x * C::WARPS + warp;
    const int groups = g.nlp >> 5;
    if (wi >= g.n_z * g.seg_count * groups) return;   // whole warps only; nothing below synchronises across warps
    const int llg = wi % groups, r = wi / groups;
    const int sidx = r % g.seg_count, z = r / g.seg_count;
    const int seg = g.seg0 + sidx;
    const int ll = llg * 32 + lane, line = g.line0 + ll;
    const int cid = (z * g.n_seg + sidx) * g.nlp + ll;
    const float4 *A, *B; short2* nnf; float* cost;
    pm_select<false>(a, z, A, B, nnf, cost);
    int start, steps;
    chain_span<DIR>(seg, seg_len, g.len, start, steps);
    if (ll >= g.n_line) steps = 0;
    int t_hi = min(steps, g.t1);
    if (g.defer && seg == 0) t_hi = min(t_hi, seg_len - 1);
    short2 prev = make_short2(0, 0);
    if (steps > 0) prev = st_prev[cid];
    // per-lane sample sites of the rounds: sample s = lane + 32 r at (i, j) = (-9 + STRIDE * (s / NJ), -9 + STRIDE * (s % NJ))
    int soff[2 * C::RP];
    float sgg[2 * C::RP];
#pragma unroll
    for (int q = 0; q < 2 * C::RP; q++) {
        int s = lane + 32 * q;
        if (s >= C::NS) s = C::NS - 1;   // lanes past the last sample score a valid site and do not store
        const int i = -PATCH_R + STRIDE * (s / C::NJ), j = -PATCH_R + STRIDE * (s % C::NJ);
        soff[q] = i * a.pw + j;
        sgg[q] = lut.gg[i < 0 ? -i : i][j < 0 ? -j : j];
    }
    const int t_max = __reduce_max_sync(0xffffffffu, t_hi);
    for (int t = g.t0; t <= t_max; t++) {
        bool need = false;
        int id = 0, pos = 0;
        short2 cur = make_short2(0, 0);
        if (t <= t_hi) {
            const int i = FWD ? start + t : start - t;
            const int x1 = ROW ? i : line, y1 = ROW ? line : i;
            id = y1 * a.w + x1;
            pos = x1 | (y1 << 16);
            if (DIR == 0) prev.x = min(prev.x + 1, a.w - 1);   // :1065/:1095/:1125/:1155
            if (DIR == 1) prev.y = min(prev.y + 1, a.h - 1);
            if (DIR == 2) prev.x = max(prev.x - 1, 0);
            if (DIR == 3) prev.y = max(prev.y - 1, 0);
            cur = nnf[id];
            // a candidate equal to the current target would be scored by the evaluation that produced cost[id]: never '<'
            need = !(prev.x == cur.x && prev.y == cur.y);
        }
        const int cand = (int)(unsigned short)prev.x | ((int)prev.y << 16);
        unsigned todo = __ballot_sync(0xffffffffu, need);
        while (todo) {
            int my_slot = -1;
#pragma unroll 1
            for (int k = 0; k < C::SLOTS && todo; k++) {
                const int e = __ffs(todo) - 1;
                todo &= todo - 1;
                if (lane == e) my_slot = k;
                const int epos = __shfl_sync(0xffffffffu, pos, e), ecand = __shfl_sync(0xffffffffu, cand, e);
                const unsigned oa = (unsigned)((epos & 0xffff) + PAD) + (unsigned)((epos >> 16) + PAD) * (unsigned)a.pw;
                const unsigned ob = (unsigned)((short)(ecand & 0xffff) + PAD) + (unsigned)((ecand >> 16) + PAD) * (unsigned)a.pw;
                const PixPk c1k = pack_pix(ldpix(A + oa)), c2k = pack_pix(ldpix(B + ob));
                float2* slot = s_val[warp][k];
#pragma unroll
                for (int q = 0; q < C::RP; q++) {
                    const float4 p1a = ldpix(A + (oa + (unsigned)soff[2 * q])), p1b = ldpix(A + (oa + (unsigned)soff[2 * q + 1]));
                    const float4 p2a = ldpix(B + (ob + (unsigned)soff[2 * q])), p2b = ldpix(B + (ob + (unsigned)soff[2 * q + 1]));
                    const PixPk p1ka = pack_pix(p1a), p1kb = pack_pix(p1b);
                    f32x2 ct, t2;
                    sample_eval2(p1a, p1ka, p1b, p1kb, p2a, p2b, c2k, c2k, pk2(max3abs_diff(c1k, p1ka), max3abs_diff(c1k, p1kb)), lut_base, ct, t2);
                    float ca, cb, wa, wb;
                    upk2(ct, ca, cb);
                    upk2(sample_weight2(t2, sgg[2 * q], sgg[2 * q + 1]), wa, wb);
                    if (lane + 64 * q < C::NS) slot[lane + 64 * q] = make_float2(ca, wa);
                    if (lane + 64 * q + 32 < C::NS) slot[lane + 64 * q + 32] = make_float2(cb, wb);
                }
            }
            __syncwarp();
            if (my_slot >= 0) {
                // the owner adds the samples of its evaluation in the reference's order (i outer, j inner; :274-296) and applies the strict '<'
                const float2* slot = s_val[warp][my_slot];
                float cs = 0.f, ws = 0.f;
#pragma unroll 4
                for (int s = 0; s < C::NS; s++) {
                    const float2 v = slot[s];
                    cs = __fmaf_rn(v.x, v.y, cs);
                    ws = __fadd_rn(ws, v.y);
                }
                const float cv = __fdiv_rn(cs, ws);
                if (cv < cost[id]) {
                    nnf[id] = prev;
                    cost[id] = cv;
                } else {
                    prev = cur;
                }
            }
            __syncwarp();
        }
    }
    if (g.defer && seg == 0 && steps > 0) st_prev[cid] = prev;   // handed to the tail launch
}

template <int DIR, int STRIDE>
static void launch_propagate_chain(eppm_context* c, const PmArgs& a, int n) {
    typedef ChainCfg<STRIDE> C;
    const bool row = (DIR == 0 || DIR == 2), fwd = DIR < 2;
    const int sl = c->prm.prop_seg_length;
    ChainGeom g;
    g.n_line = row ? a.y1 - a.y0 : a.w;
    g.line0 = row ? a.y0 : 0;
    g.nlp = (g.n_line + 31) & ~31;
    g.n_seg = row ? (a.w + sl - 1) / sl : (a.y1 + sl - 1) / sl - a.y0 / sl;
    g.seg0 = row ? 0 : a.y0 / sl;
    g.len = row ? a.w : a.h;
    g.n_z = a.n_dirs * n;
    g.seg_count = g.n_seg;
    g.t0 = 1; g.t1 = sl;
    // segment 0's last step follows segment 1's first one (both update pixel seg_len): it runs in a second launch
    g.defer = fwd && g.seg0 == 0 && (g.len + sl - 1) / sl >= 2;
    const int chains = g.n_z * g.n_seg * g.nlp;
    k_prop_snapshot<DIR><<<(chains + 255) / 256, 256, 0, c->stream>>>(a, g, sl, c->prop_prev);
    const int warps = chains / 32;
    k_prop_chain<DIR, STRIDE><<<(warps + C::WARPS - 1) / C::WARPS, C::WARPS * 32, 0, c->stream>>>(a, g, sl, c->prop_prev, c->cost_lut);
    EPPM_LAUNCH_COUNT(2);
    if (g.defer) {
        ChainGeom gt = g;
        gt.seg_count = 1;
        gt.t0 = gt.t1 = sl;
        gt.defer = 0;
        const int twarps = g.n_z * (g.nlp / 32);
        k_prop_chain<DIR, STRIDE><<<(twarps + C::WARPS - 1) / C::WARPS, C::WARPS * 32, 0, c->stream>>>(a, gt, sl, c->prop_prev, c->cost_lut);
        EPPM_LAUNCH_COUNT(1);
    }
}

// Random search (d_update_random_guess): num_guess candidates drawn in windows of radius 30,15,7,3,1,1 around the
// ENTRY best target, evaluated in order with strict '<'.
template <int STRIDE>
__global__ void __launch_bounds__(128) k_pm_search(PmArgs a, const short2* __restrict__ rng, int num_guess, int search_range, int radius_min,
                                                   const __grid_constant__ CostLut lut) {
    __shared__ float s_census[CENSUS_LUT_N];
    load_census_lut(s_census, lut);
    // CTA = blockDim.x consecutive pixels of blockDim.y consecutive rows (a 2-D tile: the patch windows of its pixels overlap in both directions)
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = a.y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= a.w || y >= a.y1) return;
    const float4 *A, *B; short2* nnf; float* cost;
    pm_select<false>(a, blockIdx.z, A, B, nnf, cost);
    const int id = y * a.w + x;
    short2 best = nnf[id];
    float best_cost = cost[id];
    const short2 entry = best;
    int mag = search_range;
    for (int k = 0; k < num_guess; k++) {
        const short2 rr = rng[(size_t)k * a.w * a.h + id];
        // :1557-1563: short sign-extended into the unsigned draw, window clipped to [0,w] x [0,h], unsigned modulo
        const unsigned r1 = (unsigned)(int)rr.x, r2 = (unsigned)(int)rr.y;
        const short xmin = (short)max(entry.x - mag, 0), xmax = (short)min(entry.x + mag + 1, a.w + 1);
        const short ymin = (short)max(entry.y - mag, 0), ymax = (short)min(entry.y + mag + 1, a.h + 1);
        const short gx = (short)(xmin + r1 % (unsigned)(xmax - xmin));
        const short gy = (short)(ymin + r2 % (unsigned)(ymax - ymin));
        if (mag / 2 >= radius_min) mag /= 2;
        const float cv = patch_cost<STRIDE, false>(A, B, a.pw, x, y, gx, gy, lut, s_census);
        if (cv < best_cost) {
            best = make_short2(gx, gy);
            best_cost = cv;
        }
    }
    nnf[id] = best;
    cost[id] = best_cost;
}

// The same search with all NG guesses of a pixel scored side by side: every guess is drawn around the ENTRY target, so the
// evaluations are independent; the image-1 side of each sample (load, range distance, spatial weight) is computed once for the NG
// candidates and the scattered image-2 gathers of the guesses overlap.  Each guess still adds its samples in the reference's order,
// and the guesses are compared in order with strict '<', so the outcome is the serial kernel's bit for bit.
// NTEX: the first NTEX guesses (the wide windows: every lane of a warp reads a different line) fetch their target-side samples
// through the texture unit instead of the LSU path, which spends one L1 wavefront per lane on them.
// NSPLIT: the NG guesses are scored in NSPLIT passes over the patch, NG / NSPLIT side by side in each: fewer live registers (centre
// colours, offsets and accumulators of the group only), more resident warps for a kernel that waits on its gathers; the image-1 side
// of a sample is recomputed per pass.
template <int STRIDE, int NG, int NTEX, int NSPLIT, int MINB>
__global__ void __launch_bounds__(MINB >= 4 ? 128 : 256, MINB) k_pm_search_joint(PmArgs a, const short2* __restrict__ rng, int search_range, int radius_min,
                                                               const __grid_constant__ CostLut lut) {
    constexpr int GS = NG / NSPLIT;
    static_assert(GS * NSPLIT == NG, "guesses must split evenly");
    __shared__ float s_census[CENSUS_LUT_N];
    load_census_lut(s_census, lut);
    // CTA = blockDim.x consecutive pixels of blockDim.y consecutive rows (a 2-D tile: the patch windows of its pixels overlap in both directions)
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = a.y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= a.w || y >= a.y1) return;
    const float4 *A, *B; short2* nnf; float* cost;
    pm_select<false>(a, blockIdx.z, A, B, nnf, cost);
    const unsigned lut_base = census_lut_base(s_census);
    const int zdir = a.n_dirs == 2 ? (blockIdx.z & 1) : 0, zb = a.n_dirs == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    const cudaTextureObject_t texB = zdir ? a.tex[0] : a.tex[1];   // a select on the block-uniform direction keeps the handle uniform
    const unsigned tb = (zdir ? a.tex_off[0] : a.tex_off[1]) + (unsigned)zb * a.plane;
    const int id = y * a.w + x;
    const short2 entry = nnf[id];
    const unsigned oa = (unsigned)(x + PAD) + (unsigned)(y + PAD) * (unsigned)a.pw;
    const PixPk c1k = pack_pix(ldpix(A + oa));
    short2 best = entry;
    float best_cost = cost[id];
    int mag = search_range;
#pragma unroll
    for (int g = 0; g < NSPLIT; g++) {
        short gx[GS], gy[GS];
        unsigned ob[GS];
        PixPk c2k[GS];
#pragma unroll
        for (int k = 0; k < GS; k++) {
            const short2 rr = rng[(size_t)(g * GS + k) * a.w * a.h + id];
            const unsigned r1 = (unsigned)(int)rr.x, r2 = (unsigned)(int)rr.y;   // :1557-1563
            const short xmin = (short)max(entry.x - mag, 0), xmax = (short)min(entry.x + mag + 1, a.w + 1);
            const short ymin = (short)max(entry.y - mag, 0), ymax = (short)min(entry.y + mag + 1, a.h + 1);
            gx[k] = (short)(xmin + r1 % (unsigned)(xmax - xmin));
            gy[k] = (short)(ymin + r2 % (unsigned)(ymax - ymin));
            if (mag / 2 >= radius_min) mag /= 2;
            ob[k] = (unsigned)(gx[k] + PAD) + (unsigned)(gy[k] + PAD) * (unsigned)a.pw;
            c2k[k] = pack_pix(ldpix(B + ob[k]));
        }
        // the guesses of a group are scored in packed pairs (sample_eval2): guess 2h in the low halves, 2h+1 in the high halves
        static_assert(GS % 2 == 0, "guesses are scored in pairs");
        f32x2 cs[GS / 2], ws[GS / 2];
#pragma unroll
        for (int h = 0; h < GS / 2; h++) cs[h] = ws[h] = pk2(0.f, 0.f);
#pragma unroll 1
        for (int i = -PATCH_R; i <= PATCH_R; i += STRIDE) {
            const int ai = i < 0 ? -i : i;
            const unsigned irow = (unsigned)(i * a.pw);
#ifndef PM_SEARCH_JUNROLL
#define PM_SEARCH_JUNROLL 1   // samples of a patch row per iteration (tuning knob; PatchMatch per pair: 1: 4.001, 2: 4.012, 5: 4.033 ms)
#endif
#define PM_PRAGMA_(x) _Pragma(#x)
#define PM_PRAGMA(x) PM_PRAGMA_(x)
PM_PRAGMA(unroll PM_SEARCH_JUNROLL)
            for (int j = -PATCH_R; j <= PATCH_R; j += STRIDE) {
                const unsigned off = irow + (unsigned)j;
                const float4 p1 = ldpix(A + (oa + off));
                const PixPk p1k = pack_pix(p1);
                const float d1s = max3abs_diff(c1k, p1k);
                const f32x2 d1 = pk2(d1s, d1s);
                const float gg = lut.gg[ai][j < 0 ? -j : j];
                f32x2 ct[GS / 2], t2[GS / 2], w[GS / 2];
#pragma unroll
                for (int h = 0; h < GS / 2; h++) {
                    const int k0 = 2 * h, k1 = 2 * h + 1;
                    const float4 p2a = g * GS + k0 < NTEX ? texpix(texB, tb + ob[k0] + off) : ldpix(B + (ob[k0] + off));
                    const float4 p2b = g * GS + k1 < NTEX ? texpix(texB, tb + ob[k1] + off) : ldpix(B + (ob[k1] + off));
                    sample_eval2(p1, p1k, p1, p1k, p2a, p2b, c2k[k0], c2k[k1], d1, lut_base, ct[h], t2[h]);
                }
                if (GS == 6) {   // one fix-up test per four + one per two samples
                    sample_weight4(t2[0], t2[1], gg, w[0], w[1]);
                    w[2] = sample_weight2(t2[2], gg, gg);
                } else {
#pragma unroll
                    for (int h = 0; h < GS / 2; h++) w[h] = sample_weight2(t2[h], gg, gg);
                }
#pragma unroll
                for (int h = 0; h < GS / 2; h++) {
                    cs[h] = fma2(ct[h], w[h], cs[h]);
                    ws[h] = add2(ws[h], w[h]);
                }
            }
        }
        // guesses are compared in their order with strict '<' (:1577); groups are visited in that order too
#pragma unroll
        for (int h = 0; h < GS / 2; h++) {
            float c0, c1, w0, w1;
            upk2(cs[h], c0, c1);
            upk2(ws[h], w0, w1);
            const float cv0 = __fdiv_rn(c0, w0), cv1 = __fdiv_rn(c1, w1);
            if (cv0 < best_cost) {
                best = make_short2(gx[2 * h], gy[2 * h]);
                best_cost = cv0;
            }
            if (cv1 < best_cost) {
                best = make_short2(gx[2 * h + 1], gy[2 * h + 1]);
                best_cost = cv1;
            }
        }
    }
    nnf[id] = best;
    cost[id] = best_cost;
}

// k_pm_search_joint at sample stride 2 on the parity-split (Q) planes: the image-1 side and the LSU-path guesses read TWO neighbouring
// samples of a patch row per 256-bit load (half the L1 requests of a kernel that is bound by them: 84 % of the L1 throughput, 46 % of the
// issue slots with 16-byte loads).  Everything else -- guesses, sample order, packed guess pairs, fix-up grouping, strict '<' -- is
// k_pm_search_joint's.
template <int NTEX, int MINB>
__global__ void __launch_bounds__(128, MINB) k_pm_search_q(PmArgs a, const short2* __restrict__ rng, int search_range, int radius_min,
                                                          const __grid_constant__ CostLut lut) {
    constexpr int NG = 6;
    __shared__ float s_census[CENSUS_LUT_N];
    load_census_lut(s_census, lut);
    // CTA = blockDim.x consecutive pixels of blockDim.y consecutive rows (a 2-D tile: the patch windows of its pixels overlap in both directions)
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = a.y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= a.w || y >= a.y1) return;
    const float4 *A, *B; short2* nnf; float* cost;
    pm_select<false>(a, blockIdx.z, A, B, nnf, cost);
    const float4 *QA, *QB;
    pm_select_q(a, blockIdx.z, QA, QB);
    const unsigned lut_base = census_lut_base(s_census);
    const int zdir = a.n_dirs == 2 ? (blockIdx.z & 1) : 0, zb = a.n_dirs == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    const cudaTextureObject_t texB = zdir ? a.tex[0] : a.tex[1];
    const unsigned tb = (zdir ? a.tex_off[0] : a.tex_off[1]) + (unsigned)zb * a.plane;
    const int id = y * a.w + x;
    const short2 entry = nnf[id];
    const PixPk c1k = pack_pix(ldpix(A + ((unsigned)(x + PAD) + (unsigned)(y + PAD) * (unsigned)a.pw)));
    unsigned ea = q_patch_origin(a.qg, x, y);
    short2 best = entry;
    float best_cost = cost[id];
    int mag = search_range;
    short gx[NG], gy[NG];
    unsigned eb[NG];      // Q origin of the guess's patch (LSU path) or, for the texture-path guesses, the packed-plane offset of its top-left sample
    PixPk c2k[NG];
#pragma unroll
    for (int k = 0; k < NG; k++) {
        const short2 rr = rng[(size_t)k * a.w * a.h + id];
        const unsigned r1 = (unsigned)(int)rr.x, r2 = (unsigned)(int)rr.y;   // :1557-1563
        const short xmin = (short)max(entry.x - mag, 0), xmax = (short)min(entry.x + mag + 1, a.w + 1);
        const short ymin = (short)max(entry.y - mag, 0), ymax = (short)min(entry.y + mag + 1, a.h + 1);
        gx[k] = (short)(xmin + r1 % (unsigned)(xmax - xmin));
        gy[k] = (short)(ymin + r2 % (unsigned)(ymax - ymin));
        if (mag / 2 >= radius_min) mag /= 2;
        const unsigned ob = (unsigned)(gx[k] + PAD) + (unsigned)(gy[k] + PAD) * (unsigned)a.pw;
        c2k[k] = pack_pix(ldpix(B + ob));
        eb[k] = k < NTEX ? tb + ob - (unsigned)(PATCH_R * a.pw + PATCH_R) : q_patch_origin(a.qg, gx[k], gy[k]);
    }
    f32x2 cs[NG / 2], ws[NG / 2];
#pragma unroll
    for (int h = 0; h < NG / 2; h++) cs[h] = ws[h] = pk2(0.f, 0.f);
#pragma unroll 1
    for (int i = 0; i < 10; i++) {
        const int ai = i < 5 ? 9 - 2 * i : 2 * i - 9;
#pragma unroll
        for (int m = 0; m < 5; m++) {
            const Pix2 P1 = ldpix2(QA + (ea + 2u * m));
            Pix2 P2[NG];
#pragma unroll
            for (int k = 0; k < NG; k++) {
                if (k < NTEX) {
                    P2[k].a = texpix(texB, eb[k] + 4u * m);
                    P2[k].b = texpix(texB, eb[k] + 4u * m + 2u);
                } else {
                    P2[k] = ldpix2(QB + (eb[k] + 2u * m));
                }
            }
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const float4 p1 = half ? P1.b : P1.a;
                const PixPk p1k = pack_pix(p1);
                const float d1s = max3abs_diff(c1k, p1k);
                const f32x2 d1 = pk2(d1s, d1s);
                const int aj = half ? (2 * m + 1 < 5 ? 7 - 4 * m : 4 * m - 7) : (2 * m < 5 ? 9 - 4 * m : 4 * m - 9);
                const float gg = lut.gg[ai][aj];
                f32x2 ct[NG / 2], t2[NG / 2], w[NG / 2];
#pragma unroll
                for (int h = 0; h < NG / 2; h++) {
                    const float4 p2a = half ? P2[2 * h].b : P2[2 * h].a, p2b = half ? P2[2 * h + 1].b : P2[2 * h + 1].a;
                    sample_eval2(p1, p1k, p1, p1k, p2a, p2b, c2k[2 * h], c2k[2 * h + 1], d1, lut_base, ct[h], t2[h]);
                }
                sample_weight4(t2[0], t2[1], gg, w[0], w[1]);
                w[2] = sample_weight2(t2[2], gg, gg);
#pragma unroll
                for (int h = 0; h < NG / 2; h++) {
                    cs[h] = fma2(ct[h], w[h], cs[h]);
                    ws[h] = add2(ws[h], w[h]);
                }
            }
        }
        ea += a.qg.qp;
#pragma unroll
        for (int k = 0; k < NG; k++) eb[k] += k < NTEX ? 2u * (unsigned)a.pw : (unsigned)a.qg.qp;
    }
#pragma unroll
    for (int h = 0; h < NG / 2; h++) {   // guesses in their order with strict '<' (:1577)
        float c0, c1, w0, w1;
        upk2(cs[h], c0, c1);
        upk2(ws[h], w0, w1);
        const float cv0 = __fdiv_rn(c0, w0), cv1 = __fdiv_rn(c1, w1);
        if (cv0 < best_cost) {
            best = make_short2(gx[2 * h], gy[2 * h]);
            best_cost = cv0;
        }
        if (cv1 < best_cost) {
            best = make_short2(gx[2 * h + 1], gy[2 * h + 1]);
            best_cost = cv1;
        }
    }
    nnf[id] = best;
    cost[id] = best_cost;
}

template <int DIR, int STRIDE>
static void launch_propagate(eppm_context* c, const PmArgs& a, int n) {
    const bool row = (DIR == 0 || DIR == 2);
    const int sl = c->prm.prop_seg_length;
    const int n_line = row ? a.y1 - a.y0 : a.w;
    // segments per line: all of them for row passes; only those inside the band for column passes (bands are segment aligned)
    const int n_seg = row ? (a.w + sl - 1) / sl : (a.y1 + sl - 1) / sl - a.y0 / sl;
    int lines = 32;  // adjacent scan lines per CTA = coalescing width; all segments of a line stay in one CTA (lock-step barrier)
    static const int env_lines = getenv("EPPM_PROP_LINES") ? atoi(getenv("EPPM_PROP_LINES")) : 0;   // tuning knob
    if (env_lines > 0) lines = env_lines;
    while (lines > 1 && lines * n_seg > 896) lines >>= 1;
    dim3 blk(lines, n_seg), grd((n_line + lines - 1) / lines, 1, a.n_dirs * n);
    if (c->variant & (EPPM_VAR_PROP_NOSKIP | EPPM_VAR_PROP_NOCOMPACT))
        k_pm_propagate<DIR, STRIDE><<<grd, blk, 0, c->stream>>>(a, c->prm.prop_seg_length, !(c->variant & EPPM_VAR_PROP_NOSKIP), c->cost_lut);
    else
        k_pm_propagate_c<DIR, STRIDE><<<grd, blk, (size_t)lines * n_seg * (sizeof(int2) + sizeof(float)), c->stream>>>(a, c->prm.prop_seg_length, c->cost_lut);
    EPPM_LAUNCH_COUNT(1);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// baoCudaPatchMatch_PlaneFitting (bao_pmflow_kernel.cu:1897-1963): the same PatchMatch with the plane-fitting cost of the refine
// stage, _d_compute_patch_dist_planefitting (:334-513) = min over the identity and three fixed affine patch models.  Declared by the
// reference's host class, called nowhere; one direction, one pair (legacy stage ABI only), plain one-thread-per-evaluation kernels.
__constant__ float c_pf_pm[3][4] = {
    {0.177f, -0.011f, -0.003f, 0.301f},   // COEF_FL_{U_X,U_Y,V_X,V_Y} (:319-332)
    {0.125f, -0.357f, 0.009f, 0.308f},    // COEF_LEFT_*
    {0.205f, 0.370f, 0.011f, 0.296f},     // COEF_RIGHT_*
};

__device__ float patch_cost_pf(const float4* __restrict__ A, const float4* __restrict__ B, int pw, int x1, int y1, int x2, int y2, const CostLut& lut,
                               const float* s_census) {
    const float4* a0 = A + (unsigned)((y1 + PAD) * pw + x1 + PAD);
    const PixPk c1k = pack_pix(ldpix(a0));
    const PixPk c2k = pack_pix(ldpix(B + (unsigned)((y2 + PAD) * pw + x2 + PAD)));
    const float uu = (float)(x2 - x1), vv = (float)(y2 - y1);   // :350-351
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, ws[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int i = -PATCH_R; i <= PATCH_R; i += 2) {
        const float fi = (float)i;
        const int ai = i < 0 ? -i : i;
        const float by = __fadd_rn((float)(y1 + i), vv);
#pragma unroll 2
        for (int j = -PATCH_R; j <= PATCH_R; j += 2) {
            const float fj = (float)j;
            const float4 p1 = ldpix(a0 + i * pw + j);
            const PixPk p1k = pack_pix(p1);
            const float d1 = max3abs_diff(c1k, p1k);
            const float gg = lut.gg[ai][j < 0 ? -j : j];
            const float bx = __fadd_rn(uu, (float)(x1 + j));
            int sx[4], sy[4];
            sx[0] = x2 + j; sy[0] = y2 + i;   // identity model: exact integers
#pragma unroll
            for (int q = 0; q < 3; q++) {     // cx2 = fma(i, C_uy, fma(j, C_ux, bx)), point fetch = floor (:402,:440,:478 as contracted)
                sx[q + 1] = __float2int_rd(__fmaf_rn(fi, c_pf_pm[q][1], __fmaf_rn(fj, c_pf_pm[q][0], bx)));
                sy[q + 1] = __float2int_rd(__fmaf_rn(fi, c_pf_pm[q][3], __fmaf_rn(fj, c_pf_pm[q][2], by)));
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
                sample_term(p1, p1k, ldpix(B + (unsigned)((sy[q] + PAD) * pw + sx[q] + PAD)), c2k, d1, gg, s_census, cs[q], ws[q]);
        }
    }
    const float k1 = __fdiv_rn(cs[0], ws[0]), k2 = __fdiv_rn(cs[1], ws[1]), k3 = __fdiv_rn(cs[2], ws[2]), k4 = __fdiv_rn(cs[3], ws[3]);
    const float m34 = k3 < k4 ? k3 : k4, m234 = k2 < m34 ? k2 : m34;   // :512 __min(cost1,__min(cost2,__min(cost3,cost4)))
    return k1 < m234 ? k1 : m234;
}

__global__ void __launch_bounds__(128) k_pf_init(PmArgs a, const short2* __restrict__ rng_init, const __grid_constant__ CostLut lut) {
    __shared__ float s_census[CENSUS_LUT_N];
    load_census_lut(s_census, lut);
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.w) return;
    const short2 t = rng_init[y * a.w + x];
    a.nnf[0][y * a.w + x] = t;
    a.cost[0][y * a.w + x] = patch_cost_pf(a.pix[0], a.pix[1], a.pw, x, y, t.x, t.y, lut, s_census);
}

// the four segment passes, CTA lock-step like k_pm_propagate (same barriers, same skip of candidates equal to the current target)
template <int DIR>
__global__ void __launch_bounds__(896) k_pf_propagate(PmArgs a, int seg_len, const __grid_constant__ CostLut lut) {
    constexpr bool ROW = (DIR == 0 || DIR == 2), FWD = (DIR < 2);
    __shared__ float s_census[CENSUS_LUT_N];
    load_census_lut(s_census, lut);
    const int line = blockIdx.x * blockDim.x + threadIdx.x, seg = threadIdx.y;
    const int n_line = ROW ? a.h : a.w, len = ROW ? a.w : a.h;
    short2* nnf = a.nnf[0];
    float* cost = a.cost[0];
    int start, steps;
    if (FWD) {
        start = seg == 0 ? 0 : seg * seg_len - 1;   // :1340-1343
        steps = min(len - 1, start + seg_len) - start;
    } else {
        start = (seg + 1) * seg_len;
        if (start >= len) start = len - 1;
        steps = start - seg * seg_len;
    }
    if (line >= n_line) steps = 0;
    auto idx = [&](int i) -> int { return ROW ? line * a.w + i : i * a.w + line; };
    short2 prev = make_short2(0, 0);
    if (steps > 0) prev = nnf[idx(start)];
    __syncthreads();
    for (int t = 1; t <= seg_len; t++) {
        if (t <= steps) {
            const int i = FWD ? start + t : start - t, id = idx(i);
            if (DIR == 0) prev.x = min(prev.x + 1, a.w - 1);
            if (DIR == 1) prev.y = min(prev.y + 1, a.h - 1);
            if (DIR == 2) prev.x = max(prev.x - 1, 0);
            if (DIR == 3) prev.y = max(prev.y - 1, 0);
            const short2 cur = nnf[id];
            if (!(prev.x == cur.x && prev.y == cur.y)) {
                const float cv = patch_cost_pf(a.pix[0], a.pix[1], a.pw, ROW ? i : line, ROW ? line : i, prev.x, prev.y, lut, s_census);
                if (cv < cost[id]) { nnf[id] = prev; cost[id] = cv; }
                else prev = cur;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128) k_pf_search(PmArgs a, const short2* __restrict__ rng, int num_guess, int search_range, int radius_min,
                                                   const __grid_constant__ CostLut lut) {
    __shared__ float s_census[CENSUS_LUT_N];
    load_census_lut(s_census, lut);
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.w) return;
    const int id = y * a.w + x;
    short2 best = a.nnf[0][id];
    float best_cost = a.cost[0][id];
    const short2 entry = best;
    int mag = search_range;
    for (int k = 0; k < num_guess; k++) {
        const short2 rr = rng[(size_t)k * a.w * a.h + id];
        const unsigned r1 = (unsigned)(int)rr.x, r2 = (unsigned)(int)rr.y;   // :1719-1726
        const short xmin = (short)max(entry.x - mag, 0), xmax = (short)min(entry.x + mag + 1, a.w + 1);
        const short ymin = (short)max(entry.y - mag, 0), ymax = (short)min(entry.y + mag + 1, a.h + 1);
        const short gx = (short)(xmin + r1 % (unsigned)(xmax - xmin)), gy = (short)(ymin + r2 % (unsigned)(ymax - ymin));
        if (mag / 2 >= radius_min) mag /= 2;
        const float cv = patch_cost_pf(a.pix[0], a.pix[1], a.pw, x, y, gx, gy, lut, s_census);
        if (cv < best_cost) { best = make_short2(gx, gy); best_cost = cv; }
    }
    a.nnf[0][id] = best;
    a.cost[0][id] = best_cost;
}

template <int DIR>
static void launch_pf_propagate(eppm_context* c, const PmArgs& a) {
    const bool row = (DIR == 0 || DIR == 2);
    const int sl = c->prm.prop_seg_length;
    const int n_line = row ? a.h : a.w, n_seg = ((row ? a.w : a.h) + sl - 1) / sl;
    int lines = 32;
    while (lines > 1 && lines * n_seg > 896) lines >>= 1;
    k_pf_propagate<DIR><<<dim3((n_line + lines - 1) / lines), dim3(lines, n_seg), 0, c->stream>>>(a, sl, c->cost_lut);
    EPPM_LAUNCH_COUNT(1);
}

// forward direction of pair 0 on the context's coarsest-level planes; patch stride 2 only (the reference's compile-time stride)
bool run_patchmatch_planefitting(eppm_context* c) {
    if (c->prm.patch_stride != 2) { set_error("plane-fitting PatchMatch is built for patch stride 2"); return false; }
    ensure_rng_tables(c);
    const int L = c->n_levels - 1;
    const LevelGeom& g = c->lv[L];
    PmArgs a = {};
    a.pix[0] = c->pix[0][L]; a.pix[1] = c->pix[1][L];
    a.plane = (unsigned)g.plane; a.pw = g.pw; a.ph = g.ph;
    a.nnf[0] = c->nnf[0]; a.cost[0] = c->cost[0];
    a.w = g.w; a.h = g.h; a.n_dirs = 1; a.y0 = 0; a.y1 = g.h;
    dim3 blk(128), grd((g.w + 127) / 128, g.h);
    k_pf_init<<<grd, blk, 0, c->stream>>>(a, c->rng_init, c->cost_lut);
    EPPM_LAUNCH_COUNT(1);
    for (int it = 0; it < c->prm.num_iter; it++) {
        launch_pf_propagate<0>(c, a);
        launch_pf_propagate<1>(c, a);
        launch_pf_propagate<2>(c, a);
        launch_pf_propagate<3>(c, a);
        k_pf_search<<<grd, blk, 0, c->stream>>>(a, c->rng_search + (size_t)it * c->prm.num_rand_guess * g.w * g.h, c->prm.num_rand_guess, c->prm.search_range,
                                                c->prm.search_radius_min, c->cost_lut);
        EPPM_LAUNCH_COUNT(1);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// baoCudaPatchMatch_Scaled (bao_pmflow_kernel.cu:1828-1895): PatchMatch over (target, patch scale).  Declared by the reference's host
// class, called nowhere, and visibly unfinished upstream -- mirrored as it stands, quirks included, so that it is bit-exact against the
// reference build on identical buffers:
//  * the cost is the bilateral AD term only (_d_compute_patch_dist_scaled :588-634; the census lines are commented out), image-2 samples
//    at (x2 + float(j)*scale, y2 + float(i)*scale) -- contracted by nvcc to fma(scale, float(j), float(x2)) (read from the reference
//    SASS) -- through a point-filtered clamped fetch = floor + replicated border;
//  * scale = float(10 + (r2 % 9) - 4) / 10.0f in UNSIGNED arithmetic = (r2 % 9 + 6) / 10 in [0.6, 1.4] (:138, :1631; the comment there
//    says 0.9~1.3), from the SECOND draw of a pixel -- the one that also gives the target's y;
//  * the forward row pass stores the candidate's SCALE into the cost plane when the candidate wins (:1207), so later comparisons at that
//    pixel run against a scale, and a candidate equal to the pixel's current (target, scale) can win: nothing is skipped here;
//  * the random field kernel indexes the scale plane with the displacement pitch (:151): the entry point requires the two pitches equal.
// One direction, one pair (legacy stage ABI only), plain one-thread-per-evaluation kernels with the CTA lock-step of k_pm_propagate.
__device__ __forceinline__ float scale_of_draw(unsigned r2) { return __fdiv_rn((float)(r2 % 9u + 6u), 10.0f); }

__global__ void k_rng_scale_tables(float* __restrict__ init, float* __restrict__ search, int w, int h, int gx, int gy, int num_iter, int num_guess,
                                   unsigned long long seed, int philox) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (philox) {   // same stream layout as k_rng_tables_philox
        if (tid >= w * h) return;
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)tid, 0, &st);
        (void)curand(&st);
        init[tid] = scale_of_draw(curand(&st));
        for (int k = 0; k < num_iter * num_guess; k++) {
            (void)curand(&st);
            search[(size_t)k * w * h + tid] = scale_of_draw(curand(&st));
        }
        return;
    }
    if (tid >= gx * gy) return;
    const int bx = tid % gx, by = tid / gx;
    curandState st;
    curand_init(seed, tid, 0, &st);
    for (int i = 0; i < RB; i++)
        for (int j = 0; j < RB; j++) {
            (void)curand(&st);
            const unsigned r2 = curand(&st);
            const int x = bx * RB + j, y = by * RB + i;
            if (x < w && y < h) init[(size_t)y * w + x] = scale_of_draw(r2);   // :138
        }
    for (int k = 0; k < num_iter * num_guess; k++)
        for (int i = 0; i < RB; i++)
            for (int j = 0; j < RB; j++) {
                (void)curand(&st);
                const unsigned r2 = curand(&st);
                const int x = bx * RB + j, y = by * RB + i;
                if (x < w && y < h) search[((size_t)k * h + y) * w + x] = scale_of_draw(r2);   // :1631
            }
}

__device__ float patch_cost_scaled(const float4* __restrict__ A, const float4* __restrict__ B, int pw, int x1, int y1, int x2, int y2, float scale,
                                   const CostLut& lut) {
    const float4* a0 = A + (unsigned)((y1 + PAD) * pw + x1 + PAD);
    const float4 c1 = ldpix(a0);
    const float4 c2 = ldpix(B + (unsigned)((y2 + PAD) * pw + x2 + PAD));
    const float fx2 = (float)x2, fy2 = (float)y2;
    float cs = 0.f, ws = 0.f;
#pragma unroll 1
    for (int i = -PATCH_R; i <= PATCH_R; i += 2) {
        const int ai = i < 0 ? -i : i;
        const int sy = __float2int_rd(__fmaf_rn(scale, (float)i, fy2));   // |i| * 1.4 = 12.6 < PAD: the replicated border is the clamp
#pragma unroll 2
        for (int j = -PATCH_R; j <= PATCH_R; j += 2) {
            const int sx = __float2int_rd(__fmaf_rn(scale, (float)j, fx2));
            const float4 p1 = ldpix(a0 + i * pw + j);
            const float4 p2 = ldpix(B + (unsigned)((sy + PAD) * pw + sx + PAD));
            const float cost = exp_ad_cost(max3abs_diff(p1, p2));                                       // :610-611
            const float d1 = max3abs_diff(c1, p1), d2 = max3abs_diff(c2, p2);
            const float e = exp_ref(div_neg_0p01(__fmaf_rn(d1, d1, __fmul_rn(d2, d2))));               // :613-617 as contracted
            const float wgt = __fmul_rn(e, lut.gg[ai][j < 0 ? -j : j]);
            cs = __fmaf_rn(cost, wgt, cs);                                                              // :619,:622 as contracted
            ws = __fadd_rn(ws, wgt);
        }
    }
    return __fdiv_rn(cs, ws);
}

struct ScaledArgs {
    const float4* pix[2];
    int pw, w, h;
    short2* nnf;
    float* scale;
    float* cost;
};

__global__ void __launch_bounds__(128) k_sc_init(ScaledArgs a, const short2* __restrict__ rng_init, const float* __restrict__ scale_init,
                                                 const __grid_constant__ CostLut lut) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.w) return;
    const int id = y * a.w + x;
    const short2 t = rng_init[id];
    const float s = scale_init[id];
    a.nnf[id] = t;
    a.scale[id] = s;
    a.cost[id] = patch_cost_scaled(a.pix[0], a.pix[1], a.pw, x, y, t.x, t.y, s, lut);   // :647-656
}

template <int DIR>
__global__ void __launch_bounds__(896) k_sc_propagate(ScaledArgs a, int seg_len, const __grid_constant__ CostLut lut) {
    constexpr bool ROW = (DIR == 0 || DIR == 2), FWD = (DIR < 2);
    const int line = blockIdx.x * blockDim.x + threadIdx.x, seg = threadIdx.y;
    const int n_line = ROW ? a.h : a.w, len = ROW ? a.w : a.h;
    int start, steps;
    if (FWD) {
        start = seg == 0 ? 0 : seg * seg_len - 1;   // :1189-1192
        steps = min(len - 1, start + seg_len) - start;
    } else {
        start = (seg + 1) * seg_len;                // :1225-1227
        if (start >= len) start = len - 1;
        steps = start - seg * seg_len;
    }
    if (line >= n_line) steps = 0;
    auto idx = [&](int i) -> int { return ROW ? line * a.w + i : i * a.w + line; };
    short2 prev = make_short2(0, 0);
    float prev_scale = 0.f;
    if (steps > 0) { prev = a.nnf[idx(start)]; prev_scale = a.scale[idx(start)]; }
    __syncthreads();
    for (int t = 1; t <= seg_len; t++) {
        if (t <= steps) {
            const int i = FWD ? start + t : start - t, id = idx(i);
            if (DIR == 0) prev.x = min(prev.x + 1, a.w - 1);
            if (DIR == 1) prev.y = min(prev.y + 1, a.h - 1);
            if (DIR == 2) prev.x = max(prev.x - 1, 0);
            if (DIR == 3) prev.y = max(prev.y - 1, 0);
            const float cv = patch_cost_scaled(a.pix[0], a.pix[1], a.pw, ROW ? i : line, ROW ? line : i, prev.x, prev.y, prev_scale, lut);
            if (cv < a.cost[id]) {
                a.nnf[id] = prev;
                a.scale[id] = prev_scale;
                a.cost[id] = DIR == 0 ? prev_scale : cv;   // :1207: the forward row pass stores the SCALE into the cost plane
            } else {
                prev = a.nnf[id];
                prev_scale = a.scale[id];
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128) k_sc_search(ScaledArgs a, const short2* __restrict__ rng, const float* __restrict__ rng_scale, int num_guess,
                                                   int search_range, int radius_min, const __grid_constant__ CostLut lut) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.w) return;
    const int id = y * a.w + x;
    short2 best = a.nnf[id];
    float best_scale = a.scale[id], best_cost = a.cost[id];
    const short2 entry = best;   // :1637-1638: every window is centred on the target the pixel entered with
    int mag = search_range;
    for (int k = 0; k < num_guess; k++) {
        const short2 rr = rng[(size_t)k * a.w * a.h + id];
        const float gs = rng_scale[(size_t)k * a.w * a.h + id];
        const unsigned r1 = (unsigned)(int)rr.x, r2 = (unsigned)(int)rr.y;   // :1635-1636: the 16-bit draws, sign-extended
        const short xmin = (short)max(entry.x - mag, 0), xmax = (short)min(entry.x + mag + 1, a.w + 1);
        const short ymin = (short)max(entry.y - mag, 0), ymax = (short)min(entry.y + mag + 1, a.h + 1);
        const short gx = (short)(xmin + r1 % (unsigned)(xmax - xmin)), gy = (short)(ymin + r2 % (unsigned)(ymax - ymin));
        if (mag / 2 >= radius_min) mag /= 2;
        const float cv = patch_cost_scaled(a.pix[0], a.pix[1], a.pw, x, y, gx, gy, gs, lut);
        if (cv < best_cost) { best = make_short2(gx, gy); best_scale = gs; best_cost = cv; }
    }
    a.nnf[id] = best;
    a.scale[id] = best_scale;
    a.cost[id] = best_cost;
}

template <int DIR>
static void launch_sc_propagate(eppm_context* c, const ScaledArgs& a) {
    const bool row = (DIR == 0 || DIR == 2);
    const int sl = c->prm.prop_seg_length;
    const int n_line = row ? a.h : a.w, n_seg = ((row ? a.w : a.h) + sl - 1) / sl;
    int lines = 32;
    while (lines > 1 && lines * n_seg > 896) lines >>= 1;
    k_sc_propagate<DIR><<<dim3((n_line + lines - 1) / lines), dim3(lines, n_seg), 0, c->stream>>>(a, sl, c->cost_lut);
    EPPM_LAUNCH_COUNT(1);
}

// forward direction of pair 0 on the context's coarsest-level planes; d_scale: dense float plane [h][w] of the caller (device memory)
bool run_patchmatch_scaled(eppm_context* c, float* d_scale) {
    if (c->prm.patch_stride != 2) { set_error("scaled PatchMatch is built for patch stride 2"); return false; }
    ensure_rng_tables(c);
    const int L = c->n_levels - 1;
    const LevelGeom& g = c->lv[L];
    if ((g.w + c->prm.prop_seg_length - 1) / c->prm.prop_seg_length > 896 || (g.h + c->prm.prop_seg_length - 1) / c->prm.prop_seg_length > 896) {
        set_error("scaled PatchMatch: more than 896 segments per scan line");
        return false;
    }
    const size_t n = (size_t)g.w * g.h, n_search = (size_t)c->prm.num_iter * c->prm.num_rand_guess;
    float* tab = nullptr;   // scale draws: [1 + iterations * guesses][h][w]; this uncalled stage allocates per call like the reference does (:1835)
    if (cudaMalloc((void**)&tab, (1 + n_search) * n * sizeof(float)) != cudaSuccess) { set_error("scaled PatchMatch: out of device memory"); return false; }
    const int gx = (g.w + RB - 1) / RB, gy = (g.h + RB - 1) / RB;
    const int philox = c->prm.rng_mode == EPPM_RNG_PHILOX;
    const int n_thr = philox ? g.w * g.h : gx * gy;
    k_rng_scale_tables<<<(n_thr + 63) / 64, 64, 0, c->stream>>>(tab, tab + n, g.w, g.h, gx, gy, c->prm.num_iter, c->prm.num_rand_guess, c->prm.seed, philox);
    EPPM_LAUNCH_COUNT(1);
    ScaledArgs a = {};
    a.pix[0] = c->pix[0][L]; a.pix[1] = c->pix[1][L];
    a.pw = g.pw; a.w = g.w; a.h = g.h;
    a.nnf = c->nnf[0]; a.cost = c->cost[0]; a.scale = d_scale;
    dim3 blk(128), grd((g.w + 127) / 128, g.h);
    k_sc_init<<<grd, blk, 0, c->stream>>>(a, c->rng_init, tab, c->cost_lut);
    EPPM_LAUNCH_COUNT(1);
    for (int it = 0; it < c->prm.num_iter; it++) {
        launch_sc_propagate<0>(c, a);   // :1327-1330: row, column, row reverse, column reverse
        launch_sc_propagate<1>(c, a);
        launch_sc_propagate<2>(c, a);
        launch_sc_propagate<3>(c, a);
        const size_t off = (size_t)it * c->prm.num_rand_guess * n;
        k_sc_search<<<grd, blk, 0, c->stream>>>(a, c->rng_search + off, tab + n + off, c->prm.num_rand_guess, c->prm.search_range, c->prm.search_radius_min,
                                                c->cost_lut);
        EPPM_LAUNCH_COUNT(1);
    }
    const cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(tab);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return false; }
    return true;
}

void run_patchmatch(eppm_context* c) { run_patchmatch_dirs(c, 2); }

template <int STRIDE>
static void run_patchmatch_t(eppm_context* c, int n_dirs, int n_steps, int first_step);

void run_patchmatch_dirs(eppm_context* c, int n_dirs, int n_steps, int first_step) {
    ensure_rng_tables(c);
    // the sample stride of the patch ("pixel skipping", bao_pmflow_kernel.cu:269,272) is a compile-time constant of the kernels
    switch (c->prm.patch_stride) {
    case 1: run_patchmatch_t<1>(c, n_dirs, n_steps, first_step); break;
    case 3: run_patchmatch_t<3>(c, n_dirs, n_steps, first_step); break;
    default: run_patchmatch_t<2>(c, n_dirs, n_steps, first_step); break;
    }
}

template <int STRIDE>
static void run_patchmatch_t(eppm_context* c, int n_dirs, int n_steps, int first_step) {
    const int L = c->n_levels - 1, n = c->n_cur;
    const LevelGeom& g = c->lv[L];
    PmArgs a;
    a.pix[0] = c->pix[0][L];
    a.pix[1] = c->pix[1][L];
    a.pixT[0] = c->pixT[0];
    a.pixT[1] = c->pixT[1];
    const bool have_q = STRIDE == 2 && c->pixQ[0] != nullptr;
    a.q[0] = have_q ? c->pixQ[0] : nullptr;
    a.q[1] = have_q ? c->pixQ[1] : nullptr;
    a.q_search = have_q && (c->variant & EPPM_VAR_PM_Q) ? 1 : 0;
    a.qg = make_qgeom(g.pw, g.ph);
    a.plane = (unsigned)g.plane;
    a.pw = g.pw; a.ph = g.ph;
    a.nnf[0] = c->nnf[0]; a.nnf[1] = c->nnf[1];
    a.cost[0] = c->cost[0]; a.cost[1] = c->cost[1];
    a.w = g.w; a.h = g.h;
    a.n_dirs = n_dirs;
    a.y0 = c->band_y0; a.y1 = c->band_y1;
    // textures over the planes (video-stream mode aliases image 2 into image 1's allocation one plane further, hence the search)
    for (int i = 0; i < 2; i++) {
        a.tex[i] = 0; a.tex_off[i] = 0;
        for (int img = 0; img < 2; img++)
            if (c->tex_pm[img] && a.pix[i] >= c->tex_pm_base[img] && a.pix[i] + (size_t)n * g.plane <= c->tex_pm_base[img] + c->tex_pm_texels) {
                a.tex[i] = c->tex_pm[img];
                a.tex_off[i] = (unsigned)(a.pix[i] - c->tex_pm_base[img]);
            }
    }
    // one thread per pixel of a row segment: 96 threads when that wastes fewer lanes than 128 (480 = 5 x 96 at the 1080p PatchMatch level)
    // The thread-per-pixel kernels (initial cost, random search) run CTAs of 16 x 16 pixels: 128 pixels of ONE row touch (128 + 18) x 19 pixels of
    // each image per guess (22 per pixel), a 16 x 16 tile (16 + 18) x (16 + 18) (4.5 per pixel) -- the windows of a resident CTA stay in the L1
    // (the search is bound by L1 / L2 traffic: 58 % L1 hit rate with row-segment CTAs).  EPPM_SEARCH_ROWS (1, 2, 4, 8): rows per CTA of 128 threads, tuning knob; 1 = the row-segment CTAs of round 1.
    // measured (16 pairs, PatchMatch ms per pair): 128 x 1: 4.67, 64 x 2: 4.48, 32 x 4: 4.37, 16 x 8: 4.32, 32 x 8: 4.28, 16 x 16: 4.24 (default)
    static const int env_rows = getenv("EPPM_SEARCH_ROWS") ? atoi(getenv("EPPM_SEARCH_ROWS")) : 16;
    static const int env_cols = getenv("EPPM_SEARCH_COLS") ? atoi(getenv("EPPM_SEARCH_COLS")) : 16;
    int ry = env_rows == 1 || env_rows == 2 || env_rows == 4 || env_rows == 16 ? env_rows : 8;
    int bx = ry == 1 ? (((g.w + 95) / 96) * 96 < ((g.w + 127) / 128) * 128 ? 96 : 128) : 128 / ry;   // 128 threads: 64 x 2, 32 x 4, 16 x 8, 8 x 16
    if (env_cols > 0 && (env_cols * ry == 128 || env_cols * ry == 256)) bx = env_cols;                // e.g. 16 x 16 or 32 x 8: 256 threads
    const bool search_256 = bx * ry == 256 && c->prm.num_rand_guess == 6 && !(c->variant & (EPPM_VAR_SEARCH_SERIAL | EPPM_VAR_SEARCH_TEX3 | EPPM_VAR_SEARCH_SPLIT3 | EPPM_VAR_SEARCH_WARP)) &&
                            !(STRIDE == 2 && a.q_search);
    if (bx * ry == 256 && !search_256) bx /= 2;   // only the default joint search has a 256-thread instantiation
    dim3 blk(bx, ry), grd((g.w + bx - 1) / bx, (a.y1 - a.y0 + ry - 1) / ry, n_dirs * n);
    // steps [first_step, n_steps): 0 = random field + cost, then per iteration 4 propagation passes and 1 random search
    int step = 0;
    auto run = [&]() { const bool r = step >= first_step && step < n_steps; step++; return r; };
    if (run()) {
        const dim3 blk_i(blk.x, blk.x * blk.y > 128 ? blk.y / 2 : blk.y);   // k_pm_init is bounded to 128 threads
        const dim3 grd_i(grd.x, (a.y1 - a.y0 + blk_i.y - 1) / blk_i.y, grd.z);
        k_pm_init<STRIDE><<<grd_i, blk_i, 0, c->stream>>>(a, c->rng_init, c->cost_lut);
        EPPM_LAUNCH_COUNT(1);
        // the evaluated-candidate memo of the propagation starts empty (-1 is no target: targets lie in [0, w] x [0, h])
        cudaMemsetAsync(c->prop_memo, 0xff, (size_t)2 * n * g.w * g.h * sizeof(int4), c->stream);
    }
    for (int it = 0; it < c->prm.num_iter && step < n_steps; it++) {
        if (c->variant & (EPPM_VAR_PROP_NOSKIP | EPPM_VAR_PROP_NOCOMPACT | EPPM_VAR_PROP_CTA)) {
            if (run()) launch_propagate<0, STRIDE>(c, a, n);
            if (run()) launch_propagate<1, STRIDE>(c, a, n);
            if (run()) launch_propagate<2, STRIDE>(c, a, n);
            if (run()) launch_propagate<3, STRIDE>(c, a, n);
        } else if (c->variant & EPPM_VAR_PROP_CHAIN) {
            if (run()) launch_propagate_chain<0, STRIDE>(c, a, n);
            if (run()) launch_propagate_chain<1, STRIDE>(c, a, n);
            if (run()) launch_propagate_chain<2, STRIDE>(c, a, n);
            if (run()) launch_propagate_chain<3, STRIDE>(c, a, n);
        } else {
            if (run()) launch_propagate_queue<0, STRIDE>(c, a, n, it * 4 + 0);
            if (run()) launch_propagate_queue<1, STRIDE>(c, a, n, it * 4 + 1);
            if (run()) launch_propagate_queue<2, STRIDE>(c, a, n, it * 4 + 2);
            if (run()) launch_propagate_queue<3, STRIDE>(c, a, n, it * 4 + 3);
        }
        if (!run()) continue;
        const short2* rng = c->rng_search + (size_t)it * c->prm.num_rand_guess * g.w * g.h;
        const bool tex_ok = a.tex[0] && a.tex[1];
        const int v = c->variant;
        const bool joint = c->prm.num_rand_guess == 6 && !(v & EPPM_VAR_SEARCH_SERIAL);
        if (joint && (v & EPPM_VAR_SEARCH_WARP)) {
            // warp-cooperative search (measured slower than one thread per pixel: 8.98 vs 5.26 ms per pair of PatchMatch): 4 pixels x 6 guesses per warp at strides 2 and 3, 1 pixel at stride 1 (361 samples per evaluation)
            constexpr int SP = STRIDE == 1 ? 1 : 4, SW = 2;
            const size_t ssmem = (size_t)SW * SP * 6 * Coop<STRIDE>::NSP * sizeof(float2);
            static bool attr_s[64] = {};
            if (c->device < 0 || c->device >= 64 || !attr_s[c->device]) {
                cudaFuncSetAttribute(k_pm_search_w<STRIDE, 6, SP, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem);
                if (c->device >= 0 && c->device < 64) attr_s[c->device] = true;
            }
            dim3 sgrd((g.w + SP * SW - 1) / (SP * SW), a.y1 - a.y0, n_dirs * n);
            k_pm_search_w<STRIDE, 6, SP, SW><<<sgrd, SW * 32, ssmem, c->stream>>>(a, rng, c->prm.search_range, c->prm.search_radius_min, c->cost_lut);
            EPPM_LAUNCH_COUNT(1);
            continue;
        }
        if (joint && STRIDE == 2 && a.q_search) {
            static const int qtex = getenv("EPPM_SEARCH_QTEX") ? atoi(getenv("EPPM_SEARCH_QTEX")) : 0;   // tuning knob: guesses on the texture path
            const int nt = tex_ok ? qtex : 0;
            if (nt >= 2) k_pm_search_q<2, 4><<<grd, blk, 0, c->stream>>>(a, rng, c->prm.search_range, c->prm.search_radius_min, c->cost_lut);
            else if (nt == 1) k_pm_search_q<1, 4><<<grd, blk, 0, c->stream>>>(a, rng, c->prm.search_range, c->prm.search_radius_min, c->cost_lut);
            else k_pm_search_q<0, 4><<<grd, blk, 0, c->stream>>>(a, rng, c->prm.search_range, c->prm.search_radius_min, c->cost_lut);
            EPPM_LAUNCH_COUNT(1);
            continue;
        }
#define EPPM_SEARCH(NT, NS, MB) k_pm_search_joint<STRIDE, 6, NT, NS, MB><<<grd, blk, pm_pad_bytes(c, (const void*)k_pm_search_joint<STRIDE, 6, NT, NS, MB>), c->stream>>>(a, rng, c->prm.search_range, c->prm.search_radius_min, c->cost_lut)
        // tuning knob for the 256-thread tiles (measured alternatives of the default's 2 texture guesses / 1 pass / 2 CTAs per SM)
        static const int s256 = getenv("EPPM_SEARCH_256MODE") ? atoi(getenv("EPPM_SEARCH_256MODE")) : 0;
        if (joint && blk.x * blk.y == 256 && ((v & EPPM_VAR_SEARCH_NOTEX) || !tex_ok)) EPPM_SEARCH(0, 1, 2);
        else if (joint && blk.x * blk.y == 256 && s256 == 1) EPPM_SEARCH(2, 3, 3);   // three passes of two guesses: 85 registers, three CTAs per SM
        else if (joint && blk.x * blk.y == 256 && s256 == 2) EPPM_SEARCH(3, 1, 2);   // three guesses through the texture unit
        else if (joint && blk.x * blk.y == 256 && s256 == 3) EPPM_SEARCH(3, 3, 3);
        else if (joint && blk.x * blk.y == 256 && s256 == 4) EPPM_SEARCH(4, 1, 2);   // four guesses through the texture unit
        else if (joint && blk.x * blk.y == 256) EPPM_SEARCH(2, 1, 2);   // 256-thread tiles: two CTAs of 128 registers per SM
        else if (joint && ((v & EPPM_VAR_SEARCH_NOTEX) || !tex_ok)) EPPM_SEARCH(0, 1, 4);
        else if (joint && (v & EPPM_VAR_SEARCH_TEX3)) EPPM_SEARCH(3, 1, 4);
        else if (joint && (v & EPPM_VAR_SEARCH_SPLIT3)) EPPM_SEARCH(2, 3, 8);
        else if (joint) EPPM_SEARCH(2, 1, 4);
#undef EPPM_SEARCH
        else
            k_pm_search<STRIDE><<<grd, blk, 0, c->stream>>>(a, rng, c->prm.num_rand_guess, c->prm.search_range, c->prm.search_radius_min, c->cost_lut);
        EPPM_LAUNCH_COUNT(1);
    }
}

}  // namespace eppm
