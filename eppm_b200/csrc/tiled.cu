// BASELINE config 4 inside the library: ONE large frame pair spatially tiled over the GPUs of a box, one process (one context) per GPU,
// halo exchange and band gathers enqueued by the library on the context's stream through NCCL (NVLink / NVSwitch).
//
// Partition and data flow are those of SURVEY.md §8e (and of eppm_b200/tiled.py, the Python reference of this schedule, which stays for the
// gloo / CPU schedule test): every rank builds the full pyramids (targets of the NNF are unbounded, prepare is 1 % of the work) and owns a band
// of coarsest-level rows aligned to the propagation segment length.  Row passes and the random search are band-local; a column pass reads ONE
// boundary row of the two NNF planes from the neighbouring band (ncclSend / ncclRecv, one group per exchange); after PatchMatch the bands of
// both NNF / cost planes are gathered on every rank (the left-right check follows arbitrary targets; one NCCL group = one fused launch per
// gather point), the tiny consistency stage runs replicated, and after every refine / smoothing step the rows each rank wrote are gathered.
// All step counts and the segment length come from the context's parameters.  Kernels take the band as a row range and are otherwise
// unchanged, so the result is bit-identical to the single-GPU run.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 -- the copy torch ships is found when torch is already loaded; EPPM_NCCL_LIB overrides):
// the library has no link-time dependency on it, and callers that never tile never load it.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "eppm_internal.h"

namespace eppm {

// the slice of nccl.h this file uses (NCCL 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0 };
struct Nccl {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static Nccl g_nccl;

static bool load_nccl() {
    if (g_nccl.handle) return true;
    const char* names[] = {getenv("EPPM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)
        if (n && !h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error(std::string("tiling needs NCCL: cannot load libnccl.so.2 (") + (dlerror() ? dlerror() : "?") + "); set EPPM_NCCL_LIB"); return false; }
    Nccl n;
    n.handle = h;
#define EPPM_SYM(field, name) *(void**)(&n.field) = dlsym(h, name); if (!n.field) { set_error(std::string("NCCL symbol missing: ") + name); return false; }
    EPPM_SYM(GetUniqueId, "ncclGetUniqueId") EPPM_SYM(CommInitRank, "ncclCommInitRank") EPPM_SYM(CommDestroy, "ncclCommDestroy")
    EPPM_SYM(GroupStart, "ncclGroupStart") EPPM_SYM(GroupEnd, "ncclGroupEnd") EPPM_SYM(Send, "ncclSend") EPPM_SYM(Recv, "ncclRecv")
    EPPM_SYM(Broadcast, "ncclBroadcast") EPPM_SYM(GetErrorString, "ncclGetErrorString")
#undef EPPM_SYM
    g_nccl = n;
    return true;
}

static bool nccl_ok(int rc, const char* what) {
    if (rc == ncclSuccess) return true;
    set_error(std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error"));
    return false;
}

// bands of the coarsest level for every rank (the arithmetic of eppm_set_band)
static void band_of(const eppm_context* c, int band, int n_bands, int* y0, int* y1) {
    const LevelGeom& gc = c->lv[c->n_levels - 1];
    const int sl = c->prm.prop_seg_length, n_seg = (gc.h + sl - 1) / sl;
    const int base = n_seg / n_bands, extra = n_seg % n_bands;
    const int s0 = band * base + (band < extra ? band : extra), s1 = s0 + base + (band < extra ? 1 : 0);
    *y0 = s0 * sl;
    *y1 = s1 * sl < gc.h ? s1 * sl : gc.h;
}
// rows of `level` that belong to rank r's band
static void level_rows_of(const eppm_context* c, int r, int world, int level, int* y0, int* y1) {
    const int L = c->n_levels - 1, sh = L - level;
    int b0, b1;
    band_of(c, r, world, &b0, &b1);
    *y0 = b0 << sh;
    *y1 = b1 >= c->lv[L].h ? c->lv[level].h : (c->lv[level].h < (b1 << sh) ? c->lv[level].h : (b1 << sh));
}

// one boundary row of both NNF planes to / from the neighbouring bands before a column pass (dir +1: forward pass, -1: reverse pass)
static bool exchange_rows(eppm_context* c, int dir) {
    const int rank = c->tile_rank, world = c->tile_world;
    const LevelGeom& gc = c->lv[c->n_levels - 1];
    ncclComm_t comm = (ncclComm_t)c->tile_comm;
    const size_t row = (size_t)gc.w * sizeof(short2);
    if (!nccl_ok(g_nccl.GroupStart(), "ncclGroupStart")) return false;
    bool ok = true;
    for (int d = 0; d < 2 && ok; d++) {
        short2* p = c->nnf[d];
        if (dir > 0) {
            if (rank + 1 < world) ok = ok && nccl_ok(g_nccl.Send(p + (size_t)(c->band_y1 - 1) * gc.w, row, ncclInt8, rank + 1, comm, c->stream), "ncclSend");
            if (rank > 0) ok = ok && nccl_ok(g_nccl.Recv(p + (size_t)(c->band_y0 - 1) * gc.w, row, ncclInt8, rank - 1, comm, c->stream), "ncclRecv");
        } else {
            if (rank > 0) ok = ok && nccl_ok(g_nccl.Send(p + (size_t)c->band_y0 * gc.w, row, ncclInt8, rank - 1, comm, c->stream), "ncclSend");
            if (rank + 1 < world) ok = ok && nccl_ok(g_nccl.Recv(p + (size_t)c->band_y1 * gc.w, row, ncclInt8, rank + 1, comm, c->stream), "ncclRecv");
        }
    }
    return nccl_ok(g_nccl.GroupEnd(), "ncclGroupEnd") && ok;
}

// every rank ends up with all rows of the planes: band r is broadcast from rank r, all planes and ranks in ONE group (one fused launch)
struct GatherPlane { void* base; size_t row_bytes; int level; };
static bool gather_bands(eppm_context* c, const GatherPlane* planes, int n_planes) {
    const int world = c->tile_world;
    ncclComm_t comm = (ncclComm_t)c->tile_comm;
    if (!nccl_ok(g_nccl.GroupStart(), "ncclGroupStart")) return false;
    bool ok = true;
    for (int k = 0; k < n_planes && ok; k++)
        for (int r = 0; r < world && ok; r++) {
            int y0, y1;
            level_rows_of(c, r, world, planes[k].level, &y0, &y1);
            if (y1 <= y0) continue;
            char* p = (char*)planes[k].base + (size_t)y0 * planes[k].row_bytes;
            ok = nccl_ok(g_nccl.Broadcast(p, p, (size_t)(y1 - y0) * planes[k].row_bytes, ncclInt8, r, comm, c->stream), "ncclBroadcast");
        }
    return nccl_ok(g_nccl.GroupEnd(), "ncclGroupEnd") && ok;
}

}  // namespace eppm

using namespace eppm;

extern "C" {

int eppm_tiled_unique_id(void* id_out) {
    if (!id_out) { set_error("eppm_tiled_unique_id: null pointer"); return EPPM_ERR_ARG; }
    if (!load_nccl()) return EPPM_ERR_STATE;
    ncclUniqueId id;
    if (!nccl_ok(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) return EPPM_ERR_CUDA;
    memcpy(id_out, &id, sizeof(id));
    return EPPM_OK;
}

int eppm_tiled_init(eppm_context* c, int rank, int world, const void* unique_id) {
    if (!c || !unique_id || world < 1 || rank < 0 || rank >= world) { set_error("eppm_tiled_init: bad argument"); return EPPM_ERR_ARG; }
    if (c->max_batch < 1) return EPPM_ERR_ARG;
    const LevelGeom& gc = c->lv[c->n_levels - 1];
    const int n_seg = (gc.h + c->prm.prop_seg_length - 1) / c->prm.prop_seg_length;
    if (world > 1 && n_seg / world < 2) { set_error("eppm_tiled_init: fewer than two propagation segments per band"); return EPPM_ERR_ARG; }
    if (!load_nccl()) return EPPM_ERR_STATE;
    if (c->tile_comm) { g_nccl.CommDestroy((ncclComm_t)c->tile_comm); c->tile_comm = nullptr; }
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != c->device) cudaSetDevice(c->device);
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_t comm = nullptr;
    const bool ok = nccl_ok(g_nccl.CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
    if (!ok) return EPPM_ERR_CUDA;
    c->tile_comm = comm;
    c->tile_rank = rank;
    c->tile_world = world;
    return EPPM_OK;
}

int eppm_tiled_shutdown(eppm_context* c) {
    if (!c) return EPPM_ERR_ARG;
    if (c->tile_comm && g_nccl.CommDestroy) {
        cudaStreamSynchronize(c->stream);
        g_nccl.CommDestroy((ncclComm_t)c->tile_comm);
    }
    c->tile_comm = nullptr;
    c->tile_world = 0;
    return EPPM_OK;
}

int eppm_compute_tiled_device(eppm_context* c, const uint8_t* d_img1, const uint8_t* d_img2, float* d_flow) {
    if (!c || !d_img1 || !d_img2 || !d_flow) { set_error("eppm_compute_tiled_device: null pointer"); return EPPM_ERR_ARG; }
    if (!c->tile_comm || c->tile_world < 1) { set_error("eppm_compute_tiled_device before eppm_tiled_init"); return EPPM_ERR_STATE; }
    if (c->prm.subpixel_final) { set_error("eppm_compute_tiled_device: subpixel_final is not combined with spatial tiling"); return EPPM_ERR_ARG; }
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != c->device) cudaSetDevice(c->device);
    const int rank = c->tile_rank, world = c->tile_world, L = c->n_levels - 1;
    const LevelGeom& gc = c->lv[L];
    bool ok = true;
    c->n_cur = 1;
    run_prepare(c, d_img1, d_img2, 1);
    band_of(c, rank, world, &c->band_y0, &c->band_y1);
    run_patchmatch_dirs(c, 2, 1, 0);   // random field + initial cost on the band
    for (int it = 0; it < c->prm.num_iter && ok; it++) {
        const int s = 1 + 5 * it;      // launch groups of an iteration: row fwd, column fwd, row rev, column rev, random search
        run_patchmatch_dirs(c, 2, s + 1, s);
        ok = ok && exchange_rows(c, +1);
        run_patchmatch_dirs(c, 2, s + 3, s + 1);
        ok = ok && exchange_rows(c, -1);
        run_patchmatch_dirs(c, 2, s + 5, s + 3);
    }
    if (ok) {
        const GatherPlane pm[4] = {{c->nnf[0], (size_t)gc.w * sizeof(short2), L}, {c->nnf[1], (size_t)gc.w * sizeof(short2), L},
                                   {c->cost[0], (size_t)gc.w * sizeof(float), L}, {c->cost[1], (size_t)gc.w * sizeof(float), L}};
        ok = gather_bands(c, pm, 4);
    }
    if (ok) {
        // the consistency stage is tiny at the coarsest level: replicated on the whole field
        c->band_y0 = 0; c->band_y1 = gc.h;
        run_consistency(c);
        band_of(c, rank, world, &c->band_y0, &c->band_y1);
        for (int level = L - 1; level >= 0 && ok; level--) {
            const size_t rb = (size_t)c->lv[level].w * sizeof(float2);
            run_c2f_step(c, level, 0, nullptr);
            const GatherPlane g0 = {c->flow_tmp, rb, level};
            ok = gather_bands(c, &g0, 1);
            run_c2f_step(c, level, 1, nullptr);
            const GatherPlane g1 = {c->flow[level], rb, level};
            ok = ok && gather_bands(c, &g1, 1);
        }
    }
    if (ok) {
        run_c2f_step(c, 0, 2, nullptr);   // final smoothing flow[0] -> flow_tmp on the band
        const GatherPlane g2 = {c->flow_tmp, (size_t)c->lv[0].w * sizeof(float2), 0};
        ok = gather_bands(c, &g2, 1);
        cudaMemcpyAsync(d_flow, c->flow_tmp, (size_t)c->lv[0].w * c->lv[0].h * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream);
    }
    c->band_y0 = 0; c->band_y1 = gc.h;
    const bool cu = cuda_ok(cudaGetLastError(), "eppm_compute_tiled_device");
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
    return ok ? (cu ? EPPM_OK : EPPM_ERR_CUDA) : EPPM_ERR_CUDA;
}

int eppm_compute_tiled_host(eppm_context* c, const uint8_t* img1, const uint8_t* img2, float* flow) {
    if (!c || !img1 || !img2) { set_error("eppm_compute_tiled_host: null pointer"); return EPPM_ERR_ARG; }
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != c->device) cudaSetDevice(c->device);
    const size_t px = (size_t)c->h * c->w;
    cudaMemcpyAsync(c->d_rgb[0], img1, px * 3, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(c->d_rgb[1], img2, px * 3, cudaMemcpyHostToDevice, c->stream);
    int rc = eppm_compute_tiled_device(c, c->d_rgb[0], c->d_rgb[1], c->d_flow_out);
    if (rc == EPPM_OK && flow) cudaMemcpyAsync(flow, c->d_flow_out, px * 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream);   // flow may be NULL on ranks that do not need it
    if (!cuda_ok(cudaStreamSynchronize(c->stream), "eppm_compute_tiled_host") && rc == EPPM_OK) rc = EPPM_ERR_CUDA;
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
