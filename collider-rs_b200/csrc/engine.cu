// libcollider_b200.so — the C ABI of include/collider_b200.h over the sm_100a kernels.
// No CPU fallback: every computing entry point launches CUDA kernels on the context's device/stream.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/collider_b200.h"
#include "kernels.cuh"
#include "sort.cuh"

using namespace cb200;

namespace {
thread_local std::string g_create_err;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t reserve(size_t n, bool keep = false, cudaStream_t st = 0) {
        if (n <= cap) return cudaSuccess;
        T* q = nullptr;
        size_t ncap = std::max(n, cap + cap / 2);
        cudaError_t e = cudaMalloc(&q, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        if (keep && p && cap) {
            e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
        if (p) cudaFree(p);
        p = q;
        cap = ncap;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

enum { ST_PRE = 0, ST_STAGE_A, ST_SCAN, ST_SCATTER, ST_CAND, ST_SOLVE, ST_TAIL, ST_COUNT };
}  // namespace

struct cb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    double cell_width = 0, padding = 0;
    std::string err;
    int sm_count = 148;

    uint32_t n = 0;
    bool have_state = false;
    DevBuf<double> col[11];  // px py w h vx vy rw rh start end pub_end
    DevBuf<double> nvel[5];  // staged new vx vy rw rh end
    DevBuf<uint8_t> kind, reit;
    DevBuf<uint32_t> group, mask;
    DevBuf<uint64_t> ids;
    DevBuf<HbRec> rec;

    GridGeom geom{0, 0};
    bool geom_forced = false;
    DevBuf<uint32_t> slots;
    DevBuf<unsigned long long> tile_state;
    DevBuf<CellEntry> entries;
    DevBuf<uint2> cand;

    DevBuf<uint2> ov_pairs;
    DevBuf<unsigned long long> ov_table;
    uint32_t n_ov = 0, ov_cap_mask = 0;
    DevBuf<uint64_t> ov_ida, ov_idb;

    DevBuf<PairEvent> events;
    DevBuf<MinKey> partial;
    uint32_t grid_a = 0, grid_p = 0;
    Counters* d_ctr = nullptr;
    StepResultDev* d_res = nullptr;  // device alias of h_res (mapped)
    StepResultDev* h_res = nullptr;  // pinned, mapped

    // last step
    bool have_step = false;
    double last_T = 0;
    StepResultDev last{};
    // sorted event cache
    DevBuf<Key128> keys_a, keys_b;
    DevBuf<uint32_t> rs_hist;
    DevBuf<EventOut> ev_out;
    bool ev_sorted = false;
    uint64_t n_ev_sorted = 0;

    uint64_t launches = 0;
    cudaEvent_t ev_t[ST_COUNT + 1] = {};
    bool timing_valid = false;
    double stage_ms_sum[ST_COUNT] = {};
    double total_ms_sum = 0;
    uint64_t timed_steps = 0;
    void* pinned_stage = nullptr;
    size_t pinned_stage_bytes = 0;
};

namespace {

int fail(cb200_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    else g_create_err = msg;
    return code;
}
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(ctx, CB200_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

StateCols cols_of(cb200_ctx* c) {
    StateCols s;
    s.px = c->col[0].p; s.py = c->col[1].p; s.w = c->col[2].p; s.h = c->col[3].p;
    s.vx = c->col[4].p; s.vy = c->col[5].p; s.rw = c->col[6].p; s.rh = c->col[7].p;
    s.start = c->col[8].p; s.end = c->col[9].p; s.pub_end = c->col[10].p;
    s.kind = c->kind.p; s.group = c->group.p; s.mask = c->mask.p; s.reit = c->reit.p;
    return s;
}

std::string flags_message(uint32_t f) {
    std::string m;
    auto add = [&m](const char* s) {
        if (!m.empty()) m += "; ";
        m += s;
    };
    if (f & kFInvalidTime) add("invalid time");
    if (f & kFEndTime) add("end time must exceed present time");
    if (f & kFCircleResize) add("circle resize velocity must maintain aspect ratio");
    if (f & kFTooSmall) add("shape width/height must be at least padding");
    if (f & kFEmptyRect) add("IndexRect contains no elements");
    if (f & kFRectSpan) add("index rect wider than the device grid period");
    if (f & kFNan) add("unexpected NaN");
    if (f & kFHighTime) add("requires time < 1e50");
    if (f & kFIdNotFound) add("hitbox id not found");
    return m;
}

// Experiment kept for A/B runs (CB200_L2_PERSIST=1): mark the HbRec table persisting in L2.  Measured on B200
// (1M hitboxes, config 3): the 79 MB persisting carve-out starves the streaming kernels (stage A 60 -> 98 us,
// scatter 52 -> 156 us) and does not speed up the solve-stage gathers, so it is OFF by default.
void set_l2_window(cb200_ctx* c) {
    const char* env = getenv("CB200_L2_PERSIST");
    if (!(env && env[0] == '1')) return;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, c->device) != cudaSuccess) return;
    if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return;
    const size_t rec_bytes = (size_t)c->n * sizeof(HbRec);
    if (rec_bytes == 0) return;
    const size_t carve = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, rec_bytes);
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = c->rec.p;
    attr.accessPolicyWindow.num_bytes = std::min<size_t>(rec_bytes, (size_t)prop.accessPolicyMaxWindowSize);
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)attr.accessPolicyWindow.num_bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
    if (getenv("CB200_VERBOSE"))
        fprintf(stderr, "[cb200] L2 persist: max %d B, window max %d B, carve %zu B, rec %zu B, hitRatio %.2f\n", prop.persistingL2CacheMaxSize,
                prop.accessPolicyMaxWindowSize, carve, rec_bytes, attr.accessPolicyWindow.hitRatio);
}

int ilog2_ceil(uint64_t v) {
    int l = 0;
    while ((1ull << l) < v) ++l;
    return l;
}

// Choose the grid period from the extent of the uploaded table (host view) unless forced.
void choose_geom(cb200_ctx* c, const cb200_state_view* v) {
    if (c->geom_forced) return;
    double minx = INFINITY, maxx = -INFINITY, miny = INFINITY, maxy = -INFINITY;
    for (uint64_t i = 0; i < v->n; ++i) {
        minx = std::min(minx, v->pos_x[i]); maxx = std::max(maxx, v->pos_x[i]);
        miny = std::min(miny, v->pos_y[i]); maxy = std::max(maxy, v->pos_y[i]);
    }
    if (!(maxx >= minx)) minx = maxx = miny = maxy = 0;
    const double cx = std::min((maxx - minx) / c->cell_width + 4.0, 1e9), cy = std::min((maxy - miny) / c->cell_width + 4.0, 1e9);
    int lw = std::max(ilog2_ceil((uint64_t)cx), 6), lh = std::max(ilog2_ceil((uint64_t)cy), 6);
    // cap the table at ~4 slots per hitbox (min 2^12, max 2^26): fold the larger axis
    const int max_bits = std::min(26, std::max(12, ilog2_ceil(std::max<uint64_t>(v->n, 1) * 4)));
    while (lw + lh > max_bits) {
        if (lw >= lh) --lw;
        else --lh;
    }
    c->geom = GridGeom{lw, lh};
}

}  // namespace

extern "C" {

int cb200_abi_version(void) { return CB200_ABI_VERSION; }

const char* cb200_last_error(const cb200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int cb200_create(double cell_width, double padding, int device, cb200_ctx** out) {
    cb200_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, CB200_E_ARG, "out is null");
    *out = nullptr;
    // Collider::new (collider.rs:59-69)
    if (!(cell_width > padding)) return fail(nullptr, CB200_E_ARG, "requires cell_width > padding");
    if (!(padding > 0.0)) return fail(nullptr, CB200_E_ARG, "requires padding > 0.0");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(nullptr, CB200_E_CUDA, std::string("no usable CUDA device (no CPU fallback exists): ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, CB200_E_ARG, "device ordinal out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CB200_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cb200_ctx* c = new cb200_ctx;
    c->device = device;
    c->cell_width = cell_width;
    c->padding = padding;
    auto bail = [&](const char* what, cudaError_t err) {
        std::string m = std::string(what) + ": " + cudaGetErrorString(err);
        delete c;
        return fail(nullptr, CB200_E_CUDA, m);
    };
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return bail("cudaDeviceGetAttribute", e);
    if ((e = cudaMalloc(&c->d_ctr, sizeof(Counters))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaHostAlloc(&c->h_res, sizeof(StepResultDev), cudaHostAllocMapped)) != cudaSuccess) return bail("cudaHostAlloc", e);
    if ((e = cudaHostGetDevicePointer((void**)&c->d_res, c->h_res, 0)) != cudaSuccess) return bail("cudaHostGetDevicePointer", e);
    for (auto& ev : c->ev_t)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    *out = c;
    (void)ctx;
    return CB200_OK;
}

void cb200_destroy(cb200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& b : c->col) b.release();
    for (auto& b : c->nvel) b.release();
    c->kind.release(); c->reit.release(); c->group.release(); c->mask.release(); c->ids.release(); c->rec.release();
    c->slots.release(); c->tile_state.release(); c->entries.release();
    c->ov_pairs.release(); c->ov_table.release(); c->ov_ida.release(); c->ov_idb.release();
    c->events.release(); c->partial.release();
    c->keys_a.release(); c->keys_b.release(); c->rs_hist.release(); c->ev_out.release(); c->cand.release();
    if (c->d_ctr) cudaFree(c->d_ctr);
    if (c->h_res) cudaFreeHost(c->h_res);
    if (c->pinned_stage) cudaFreeHost(c->pinned_stage);
    for (auto& ev : c->ev_t)
        if (ev) cudaEventDestroy(ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int cb200_set_grid(cb200_ctx* ctx, int log2_w, int log2_h) {
    if (!ctx) return CB200_E_ARG;
    if (log2_w == 0 && log2_h == 0) {
        ctx->geom_forced = false;
        return CB200_OK;
    }
    if (log2_w < 1 || log2_h < 1 || log2_w > 20 || log2_h > 20 || log2_w + log2_h < 12 || log2_w + log2_h > 28)
        return fail(ctx, CB200_E_ARG, "grid period out of range (need 12 <= log2_w + log2_h <= 28)");
    ctx->geom = GridGeom{log2_w, log2_h};
    ctx->geom_forced = true;
    return CB200_OK;
}

int cb200_get_grid(cb200_ctx* ctx, int* log2_w, int* log2_h) {
    if (!ctx) return CB200_E_ARG;
    if (log2_w) *log2_w = ctx->geom.lw;
    if (log2_h) *log2_h = ctx->geom.lh;
    return CB200_OK;
}

void* cb200_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void cb200_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int cb200_upload_state(cb200_ctx* ctx, const cb200_state_view* v) {
    if (!ctx || !v) return CB200_E_ARG;
    if (v->n >= 0x7fffffffull) return fail(ctx, CB200_E_ARG, "too many hitboxes for one device (n < 2^31 - 1)");
    const void* need[] = {v->id, v->pos_x, v->pos_y, v->dim_x, v->dim_y, v->vel_x, v->vel_y, v->rsz_x, v->rsz_y,
                          v->start_time, v->end_time, v->pub_end_time, v->kind, v->group, v->interact_mask};
    if (v->n)
        for (const void* p : need)
            if (!p) return fail(ctx, CB200_E_ARG, "null column in state view");
    for (uint64_t i = 1; i < v->n; ++i)
        if (!(v->id[i - 1] < v->id[i])) return fail(ctx, CB200_E_ARG, "ids must be strictly ascending");
    for (uint64_t i = 0; i < v->n; ++i)
        if (v->group[i] != CB200_GROUP_NONE && v->group[i] > 31) return fail(ctx, CB200_E_ARG, "group must be 0..31 or CB200_GROUP_NONE");
    CU(cudaSetDevice(ctx->device));
    const size_t n = (size_t)v->n, cap = std::max<size_t>(n, 1);
    const double* src[11] = {v->pos_x, v->pos_y, v->dim_x, v->dim_y, v->vel_x, v->vel_y, v->rsz_x, v->rsz_y, v->start_time, v->end_time, v->pub_end_time};
    for (int k = 0; k < 11; ++k) {
        CU(ctx->col[k].reserve(cap));
        if (n) CU(cudaMemcpyAsync(ctx->col[k].p, src[k], n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(ctx->kind.reserve(cap)); CU(ctx->reit.reserve(cap)); CU(ctx->group.reserve(cap)); CU(ctx->mask.reserve(cap));
    CU(ctx->ids.reserve(cap)); CU(ctx->rec.reserve(cap));
    if (n) {
        CU(cudaMemcpyAsync(ctx->kind.p, v->kind, n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->group.p, v->group, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->mask.p, v->interact_mask, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->ids.p, v->id, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->reit.p, 0, n, ctx->stream));
    }
    choose_geom(ctx, v);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->n = (uint32_t)n;
    set_l2_window(ctx);
    ctx->have_state = true;
    ctx->have_step = false;
    ctx->ev_sorted = false;
    ctx->n_ov = 0;
    ctx->ov_cap_mask = 0;
    return CB200_OK;
}

int cb200_download_state(cb200_ctx* ctx, const cb200_state_out* o) {
    if (!ctx || !o) return CB200_E_ARG;
    if (!ctx->have_state) return fail(ctx, CB200_E_STATE, "no state uploaded");
    if (o->n < ctx->n) return fail(ctx, CB200_E_ARG, "state_out capacity too small");
    CU(cudaSetDevice(ctx->device));
    double* dst[11] = {o->pos_x, o->pos_y, o->dim_x, o->dim_y, o->vel_x, o->vel_y, o->rsz_x, o->rsz_y, o->start_time, o->end_time, o->pub_end_time};
    for (int k = 0; k < 11; ++k)
        if (dst[k] && ctx->n) CU(cudaMemcpyAsync(dst[k], ctx->col[k].p, (size_t)ctx->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return CB200_OK;
}

int cb200_set_overlaps(cb200_ctx* ctx, const uint64_t* id_a, const uint64_t* id_b, uint64_t n_pairs) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->have_state) return fail(ctx, CB200_E_STATE, "set_overlaps before upload_state");
    if (n_pairs && (!id_a || !id_b)) return fail(ctx, CB200_E_ARG, "null id array");
    if (n_pairs >= 0x40000000ull) return fail(ctx, CB200_E_ARG, "too many overlapping pairs");
    CU(cudaSetDevice(ctx->device));
    ctx->n_ov = 0;
    ctx->ov_cap_mask = 0;
    if (n_pairs == 0) return CB200_OK;
    const uint32_t np = (uint32_t)n_pairs;
    CU(ctx->ov_ida.reserve(np)); CU(ctx->ov_idb.reserve(np)); CU(ctx->ov_pairs.reserve(np));
    uint32_t cap = 1024;
    while (cap < 2 * np) cap <<= 1;
    CU(ctx->ov_table.reserve(cap));
    CU(cudaMemcpyAsync(ctx->ov_ida.p, id_a, (size_t)np * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->ov_idb.p, id_b, (size_t)np * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->ov_table.p, 0xff, (size_t)cap * 8, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(Counters), ctx->stream));
    const unsigned long long big = ~0ull;
    CU(cudaMemcpyAsync(&ctx->d_ctr->err_slot, &big, 8, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t blocks = (np + kThreads - 1) / kThreads;
    k_overlap_map_ids<<<blocks, kThreads, 0, ctx->stream>>>(ctx->ids.p, ctx->n, ctx->ov_ida.p, ctx->ov_idb.p, np, ctx->ov_pairs.p, ctx->d_ctr);
    k_overlap_insert<<<blocks, kThreads, 0, ctx->stream>>>(ctx->ov_pairs.p, np, ctx->ov_table.p, cap - 1);
    ctx->launches += 2;
    Counters h;
    CU(cudaMemcpyAsync(&h, ctx->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    if (h.err_flags) return fail(ctx, CB200_E_ARG, "set_overlaps: hitbox id not found (pair " + std::to_string(h.err_slot) + ")");
    ctx->n_ov = np;
    ctx->ov_cap_mask = cap - 1;
    return CB200_OK;
}

static int run_step(cb200_ctx* ctx, double time, bool rebase, const double* const* nv, double end_time);

int cb200_bulk_update(cb200_ctx* ctx, double time, const cb200_vel_view* vel, double end_time, cb200_step_result* out) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->have_state) return fail(ctx, CB200_E_STATE, "bulk_update before upload_state");
    if (!(time < 1e50)) return fail(ctx, CB200_E_ARG, "requires time < 1e50");  // collider.rs:100
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->n;
    CU(cudaEventRecord(ctx->ev_t[ST_PRE], ctx->stream));
    const double* nv[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (vel && n) {
        const double* src[5] = {vel->vel_x, vel->vel_y, vel->rsz_x, vel->rsz_y, vel->end_time};
        for (int k = 0; k < 5; ++k)
            if (src[k]) {
                CU(ctx->nvel[k].reserve(n));
                CU(cudaMemcpyAsync(ctx->nvel[k].p, src[k], (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
                nv[k] = ctx->nvel[k].p;
            }
    }
    int rc = run_step(ctx, time, true, nv, end_time);
    if (rc != CB200_OK) return rc;
    const StepResultDev& r = ctx->last;
    if (out) {
        memset(out, 0, sizeof(*out));
        out->next_time = r.next_time;
        out->next_event.time = r.next_time;
        if (r.next_tie != kKeyMax) {
            // the two ids need one tiny D2H each; cheap and only when the caller asked for the result struct
            if (r.next_tie & kPairClass) {
                const uint32_t hi = (uint32_t)((r.next_tie >> 32) & 0x7fffffffu), lo = (uint32_t)(r.next_tie & 0x7fffffffu);
                CU(cudaMemcpyAsync(&out->next_event.a, ctx->ids.p + hi, 8, cudaMemcpyDeviceToHost, ctx->stream));
                CU(cudaMemcpyAsync(&out->next_event.b, ctx->ids.p + lo, 8, cudaMemcpyDeviceToHost, ctx->stream));
                out->next_event.kind = (r.next_tie & kCollideBit) ? CB200_EV_COLLIDE : CB200_EV_SEPARATE;
            } else {
                CU(cudaMemcpyAsync(&out->next_event.a, ctx->ids.p + (uint32_t)r.next_tie, 8, cudaMemcpyDeviceToHost, ctx->stream));
                out->next_event.kind = CB200_EV_REITERATE;
            }
            CU(cudaStreamSynchronize(ctx->stream));
        }
        out->n_hitboxes = n;
        out->n_entries = r.ctr.n_entries;
        out->n_cells = r.ctr.n_cells;
        out->n_collide_solves = r.ctr.n_collide;
        out->n_separate_solves = r.ctr.n_separate;
        out->n_solitaire = r.ctr.n_solitaire;
        out->n_events = r.ctr.n_pair_events + r.ctr.n_solitaire;
        out->error_flags = r.ctr.err_flags;
        out->error_slot = 0;
    }
    if (r.ctr.err_flags) {
        uint64_t bad_id = 0;
        if (r.ctr.err_slot < n) {
            CU(cudaMemcpy(&bad_id, ctx->ids.p + r.ctr.err_slot, 8, cudaMemcpyDeviceToHost));
        }
        if (out) out->error_slot = bad_id;
        return fail(ctx, CB200_E_CONTRACT, flags_message(r.ctr.err_flags) + " (hitbox id " + std::to_string(bad_id) + ")");
    }
    return CB200_OK;
}

// One pass of the pipeline.  rebase=false re-runs binning + pairs on the already-rebased table (used after a
// buffer grew); stage A is idempotent at the same T except for the (unchanged) velocity install.
static int run_step(cb200_ctx* ctx, double time, bool rebase, const double* const* nv, double end_time) {
    const uint32_t n = ctx->n;
    const GridGeom g = ctx->geom;
    const uint32_t n_slots = g.n_slots();
    const uint32_t n_tiles = n_slots / kScanTile;
    CU(ctx->slots.reserve(n_slots));
    CU(ctx->tile_state.reserve(n_tiles));
    if (ctx->entries.cap == 0) CU(ctx->entries.reserve((size_t)n * 4 + 1024));
    if (ctx->cand.cap == 0) CU(ctx->cand.reserve((size_t)n * 2 + 1024));
    if (ctx->events.cap == 0) CU(ctx->events.reserve((size_t)n * 2 + 1024));
    const uint32_t per_sm = 8;
    ctx->grid_a = std::max<uint32_t>(1, std::min<uint32_t>((n + kThreads - 1) / kThreads, ctx->sm_count * per_sm));
    ctx->grid_p = ctx->sm_count * per_sm;
    const uint32_t n_partial = ctx->grid_a + ctx->grid_p;
    CU(ctx->partial.reserve(n_partial));

    for (int attempt = 0; attempt < 4; ++attempt) {
        const bool do_rebase = rebase && attempt == 0;
        CU(cudaMemsetAsync(ctx->slots.p, 0, (size_t)n_slots * 4, ctx->stream));
        CU(cudaMemsetAsync(ctx->tile_state.p, 0, (size_t)n_tiles * 8, ctx->stream));
        CU(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(Counters), ctx->stream));
        CU(cudaMemsetAsync(&ctx->d_ctr->err_slot, 0xff, 8, ctx->stream));
        CU(cudaEventRecord(ctx->ev_t[ST_STAGE_A], ctx->stream));

        StepArgs a;
        a.n = n;
        a.rebase = do_rebase ? 1 : 0;
        a.T = time;
        a.cell_width = ctx->cell_width;
        a.padding = ctx->padding;
        a.end_scalar = end_time;
        a.s = cols_of(ctx);
        a.nvx = do_rebase ? nv[0] : nullptr; a.nvy = do_rebase ? nv[1] : nullptr;
        a.nrw = do_rebase ? nv[2] : nullptr; a.nrh = do_rebase ? nv[3] : nullptr;
        a.nend = do_rebase ? nv[4] : ctx->col[10].p;  // re-run: the user end_time is the stored pub_end_time
        a.rec = ctx->rec.p;
        a.slot_count = ctx->slots.p;
        a.geom = g;
        a.ctr = ctx->d_ctr;
        a.partial = ctx->partial.p;
        if (n) k_stage_a<<<ctx->grid_a, kThreads, 0, ctx->stream>>>(a);
        else CU(cudaMemsetAsync(ctx->partial.p, 0xff, sizeof(MinKey) * ctx->grid_a, ctx->stream));
        CU(cudaEventRecord(ctx->ev_t[ST_SCAN], ctx->stream));
        k_scan_slots<<<n_tiles, kThreads, 0, ctx->stream>>>(ctx->slots.p, n_slots, ctx->tile_state.p, ctx->d_ctr);
        CU(cudaEventRecord(ctx->ev_t[ST_SCATTER], ctx->stream));
        if (n) k_scatter<<<ctx->grid_a, kThreads, 0, ctx->stream>>>(n, ctx->rec.p, ctx->slots.p, ctx->entries.p, (uint32_t)std::min<size_t>(ctx->entries.cap, 0xffffffffu), g, ctx->d_ctr);
        CU(cudaEventRecord(ctx->ev_t[ST_CAND], ctx->stream));
        CandArgs cd;
        cd.entries = ctx->entries.p;
        cd.geom = g;
        cd.ov = OverlapSet{ctx->ov_table.p, ctx->n_ov ? ctx->ov_cap_mask : 0u};
        cd.cand = ctx->cand.p;
        cd.cap_cand = (uint32_t)std::min<size_t>(ctx->cand.cap, 0xffffffffu);
        cd.ctr = ctx->d_ctr;
        k_candidates<<<ctx->grid_p, kThreads, 0, ctx->stream>>>(cd);
        CU(cudaEventRecord(ctx->ev_t[ST_SOLVE], ctx->stream));
        SolveArgs sv;
        sv.cand = ctx->cand.p;
        sv.cap_cand = cd.cap_cand;
        sv.ov_pairs = ctx->ov_pairs.p;
        sv.n_ov = ctx->n_ov;
        sv.rec = ctx->rec.p;
        sv.T = time;
        sv.padding = ctx->padding;
        sv.events = ctx->events.p;
        sv.cap_events = (uint32_t)std::min<size_t>(ctx->events.cap, 0xffffffffu);
        sv.ctr = ctx->d_ctr;
        sv.partial = ctx->partial.p;
        sv.n_partial_a = ctx->grid_a;
        sv.result = ctx->d_res;
        k_solve<<<ctx->grid_p, kThreads, 0, ctx->stream>>>(sv);
        CU(cudaEventRecord(ctx->ev_t[ST_TAIL], ctx->stream));
        CU(cudaEventRecord(ctx->ev_t[ST_COUNT], ctx->stream));
        ctx->launches += (n ? 2 : 0) + 3;
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaGetLastError());
        ctx->last = *ctx->h_res;
        ctx->last_T = time;
        ctx->have_step = true;
        ctx->ev_sorted = false;
        const Counters& c = ctx->last.ctr;
        bool again = false;
        if (c.overflow_entries || c.n_entries > ctx->entries.cap) {
            CU(ctx->entries.reserve((size_t)c.n_entries + c.n_entries / 4 + 1024));
            again = true;
        }
        if (c.n_collide > ctx->cand.cap) {
            CU(ctx->cand.reserve((size_t)c.n_collide + c.n_collide / 4 + 1024));
            again = true;
        }
        if (c.n_pair_events > ctx->events.cap) {
            CU(ctx->events.reserve((size_t)c.n_pair_events + c.n_pair_events / 4 + 1024));
            again = true;
        }
        if (!again) {
            float ms = 0;
            for (int s = 0; s < ST_COUNT; ++s) {
                if (cudaEventElapsedTime(&ms, ctx->ev_t[s], ctx->ev_t[s + 1]) == cudaSuccess) ctx->stage_ms_sum[s] += ms;
            }
            if (cudaEventElapsedTime(&ms, ctx->ev_t[ST_PRE], ctx->ev_t[ST_COUNT]) == cudaSuccess) ctx->total_ms_sum += ms;
            ctx->timed_steps += 1;
            ctx->timing_valid = true;
            return CB200_OK;
        }
        if (c.err_flags) return CB200_OK;  // reported by the caller
    }
    return fail(ctx, CB200_E_STATE, "step buffers failed to converge");
}

// ---- event queue materialisation -------------------------------------------------------------------------
static int sort_events(cb200_ctx* ctx) {
    if (ctx->ev_sorted) return CB200_OK;
    const Counters& c = ctx->last.ctr;
    const uint64_t n_pairs = c.n_pair_events, total = n_pairs + c.n_solitaire;
    ctx->n_ev_sorted = total;
    if (total == 0) {
        ctx->ev_sorted = true;
        return CB200_OK;
    }
    if (total >= 0xffffffffull) return fail(ctx, CB200_E_STATE, "too many events to sort");
    const uint32_t nt = (uint32_t)total;
    CU(ctx->keys_a.reserve(nt)); CU(ctx->keys_b.reserve(nt)); CU(ctx->ev_out.reserve(nt));
    const uint32_t n_blocks = (nt + kRsTile - 1) / kRsTile;
    CU(ctx->rs_hist.reserve((size_t)16 * n_blocks));
    CU(cudaMemsetAsync(&ctx->d_ctr->n_sort_keys, 0, 8, ctx->stream));
    CU(cudaMemsetAsync(&ctx->d_ctr->diff_hi, 0, 16, ctx->stream));
    if (ctx->n) k_keys_solitaire<<<(ctx->n + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(ctx->n, ctx->reit.p, ctx->col[9].p, ctx->keys_a.p, ctx->d_ctr);
    if (n_pairs) k_keys_pairs<<<((uint32_t)n_pairs + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>((uint32_t)n_pairs, ctx->events.p, ctx->keys_a.p, c.n_solitaire);
    k_key_diff<<<std::min<uint32_t>(n_blocks, 1024), kThreads, 0, ctx->stream>>>(nt, ctx->keys_a.p, ctx->d_ctr);
    ctx->launches += 3;
    unsigned long long diff[2];
    CU(cudaMemcpyAsync(diff, &ctx->d_ctr->diff_hi, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const unsigned long long diff_hi = diff[0], diff_lo = diff[1];
    Key128 *in = ctx->keys_a.p, *outp = ctx->keys_b.p;
    for (int shift = 0; shift < 128; shift += 4) {
        const unsigned long long d = shift < 64 ? (diff_lo >> shift) & 15ull : (diff_hi >> (shift - 64)) & 15ull;
        if (!d) continue;
        k_rs_hist<<<n_blocks, kRsThreads, 0, ctx->stream>>>(in, nt, shift, ctx->rs_hist.p, n_blocks);
        k_rs_scan<<<1, 1024, 0, ctx->stream>>>(ctx->rs_hist.p, 16 * n_blocks);
        k_rs_scatter<<<n_blocks, kRsThreads, 0, ctx->stream>>>(in, outp, nt, shift, ctx->rs_hist.p, n_blocks);
        ctx->launches += 3;
        std::swap(in, outp);
    }
    k_decode_events<<<(nt + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(nt, in, ctx->ids.p, ctx->ev_out.p);
    ctx->launches += 1;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    ctx->ev_sorted = true;
    return CB200_OK;
}

int cb200_fetch_events(cb200_ctx* ctx, cb200_event* buf, uint64_t cap, uint64_t offset, uint64_t* n_out) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->have_step) return fail(ctx, CB200_E_STATE, "fetch_events before bulk_update");
    CU(cudaSetDevice(ctx->device));
    static_assert(sizeof(cb200_event) == sizeof(EventOut), "event layout");
    int rc = sort_events(ctx);
    if (rc != CB200_OK) return rc;
    uint64_t avail = ctx->n_ev_sorted > offset ? ctx->n_ev_sorted - offset : 0;
    uint64_t take = std::min(avail, cap);
    if (take && !buf) return fail(ctx, CB200_E_ARG, "null event buffer");
    if (take) CU(cudaMemcpy(buf, ctx->ev_out.p + offset, take * sizeof(EventOut), cudaMemcpyDeviceToHost));
    if (n_out) *n_out = take;
    return CB200_OK;
}

int cb200_fetch_candidate_pairs(cb200_ctx* ctx, uint64_t* id_hi, uint64_t* id_lo, uint64_t cap, uint64_t* n_out) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->have_step) return fail(ctx, CB200_E_STATE, "fetch_candidate_pairs before bulk_update");
    CU(cudaSetDevice(ctx->device));
    const uint64_t total = ctx->last.ctr.n_collide;
    if (n_out) *n_out = total;
    if (cap == 0 || total == 0) return CB200_OK;
    std::vector<uint2> h(total);
    std::vector<uint64_t> ids(ctx->n);
    CU(cudaMemcpyAsync(h.data(), ctx->cand.p, total * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ids.data(), ctx->ids.p, (size_t)ctx->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (uint64_t i = 0; i < total && i < cap; ++i) {
        if (id_hi) id_hi[i] = ids[h[i].x];
        if (id_lo) id_lo[i] = ids[h[i].y];
    }
    return CB200_OK;
}

int cb200_fetch_cell_entries(cb200_ctx* ctx, int32_t* cell_x, int32_t* cell_y, uint64_t* id, uint64_t cap, uint64_t* n_out) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->have_step) return fail(ctx, CB200_E_STATE, "fetch_cell_entries before bulk_update");
    CU(cudaSetDevice(ctx->device));
    const uint64_t total = ctx->last.ctr.n_entries;
    if (n_out) *n_out = total;
    if (cap == 0 || total == 0) return CB200_OK;
    std::vector<CellEntry> h(total);
    std::vector<uint64_t> ids(ctx->n);
    CU(cudaMemcpy(h.data(), ctx->entries.p, total * sizeof(CellEntry), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ids.data(), ctx->ids.p, (size_t)ctx->n * 8, cudaMemcpyDeviceToHost));
    const uint32_t mw = (1u << ctx->geom.lw) - 1u, mh = (1u << ctx->geom.lh) - 1u;
    for (uint64_t i = 0; i < total && i < cap; ++i) {
        const CellEntry& r = h[i];
        const uint32_t sxs = r.slot & mw, sys = r.slot >> ctx->geom.lw;
        // the unique cell of the rect that folds onto this slot
        const int32_t x = r.sx + (int32_t)((sxs - (uint32_t)r.sx) & mw), y = r.sy + (int32_t)((sys - (uint32_t)r.sy) & mh);
        if (cell_x) cell_x[i] = x;
        if (cell_y) cell_y[i] = y;
        if (id) id[i] = ids[r.idx];
    }
    return CB200_OK;
}

int cb200_fetch_index_rects(cb200_ctx* ctx, int32_t* rects, uint64_t cap_hitboxes) {
    if (!ctx || !rects) return CB200_E_ARG;
    if (!ctx->have_step) return fail(ctx, CB200_E_STATE, "fetch_index_rects before bulk_update");
    if (cap_hitboxes < ctx->n) return fail(ctx, CB200_E_ARG, "rect buffer too small");
    CU(cudaSetDevice(ctx->device));
    // strided D2H: first 16 bytes of every 96-byte record
    if (ctx->n) CU(cudaMemcpy2D(rects, 16, ctx->rec.p, sizeof(HbRec), 16, ctx->n, cudaMemcpyDeviceToHost));
    return CB200_OK;
}

static int narrow_batch(cb200_ctx* ctx, uint64_t n, const double* a, const uint8_t* ka, const double* b, const uint8_t* kb, double padding,
                        bool collide, double* out) {
    if (!ctx) return CB200_E_ARG;
    if (n == 0) return CB200_OK;
    if (!a || !ka || !b || !kb || !out) return fail(ctx, CB200_E_ARG, "null array");
    CU(cudaSetDevice(ctx->device));
    double *da = nullptr, *db = nullptr, *dout = nullptr;
    uint8_t *dka = nullptr, *dkb = nullptr;
    cudaError_t e = cudaSuccess;
    auto done = [&]() {
        cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(dka); cudaFree(dkb);
    };
#define CUX(call)                                                                                  \
    do {                                                                                           \
        e = (call);                                                                                \
        if (e != cudaSuccess) {                                                                    \
            done();                                                                                \
            return fail(ctx, CB200_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e));     \
        }                                                                                          \
    } while (0)
    CUX(cudaMalloc(&da, n * 72)); CUX(cudaMalloc(&db, n * 72)); CUX(cudaMalloc(&dout, n * 8));
    CUX(cudaMalloc(&dka, n)); CUX(cudaMalloc(&dkb, n));
    CUX(cudaMemcpyAsync(da, a, n * 72, cudaMemcpyHostToDevice, ctx->stream));
    CUX(cudaMemcpyAsync(db, b, n * 72, cudaMemcpyHostToDevice, ctx->stream));
    CUX(cudaMemcpyAsync(dka, ka, n, cudaMemcpyHostToDevice, ctx->stream));
    CUX(cudaMemcpyAsync(dkb, kb, n, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t blocks = (uint32_t)((n + 255) / 256);
    if (collide) k_collide_batch<<<blocks, 256, 0, ctx->stream>>>(n, da, dka, db, dkb, dout);
    else k_separate_batch<<<blocks, 256, 0, ctx->stream>>>(n, da, dka, db, dkb, padding, dout);
    ctx->launches += 1;
    CUX(cudaMemcpyAsync(out, dout, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUX(cudaStreamSynchronize(ctx->stream));
    CUX(cudaGetLastError());
#undef CUX
    done();
    return CB200_OK;
}

int cb200_collide_time_batch(cb200_ctx* ctx, uint64_t n, const double* a, const uint8_t* kind_a, const double* b, const uint8_t* kind_b, double* out) {
    return narrow_batch(ctx, n, a, kind_a, b, kind_b, 0.0, true, out);
}
int cb200_separate_time_batch(cb200_ctx* ctx, uint64_t n, const double* a, const uint8_t* kind_a, const double* b, const uint8_t* kind_b,
                              double padding, double* out) {
    return narrow_batch(ctx, n, a, kind_a, b, kind_b, padding, false, out);
}

uint64_t cb200_launch_count(const cb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int cb200_last_step_timing(cb200_ctx* ctx, float* total_ms, float* stage_ms, int n_stages) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->timing_valid || ctx->timed_steps == 0) return fail(ctx, CB200_E_STATE, "no timed step yet");
    const double k = 1.0 / (double)ctx->timed_steps;
    if (total_ms) *total_ms = (float)(ctx->total_ms_sum * k);
    for (int s = 0; s < n_stages && s < ST_COUNT; ++s)
        if (stage_ms) stage_ms[s] = (float)(ctx->stage_ms_sum[s] * k);
    return CB200_OK;
}
int cb200_reset_timing(cb200_ctx* ctx) {
    if (!ctx) return CB200_E_ARG;
    for (double& v : ctx->stage_ms_sum) v = 0;
    ctx->total_ms_sum = 0;
    ctx->timed_steps = 0;
    ctx->timing_valid = false;
    return CB200_OK;
}

}  // extern "C"
