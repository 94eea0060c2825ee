// libcollider_b200.so — the C ABI of include/collider_b200.h over the sm_100a kernels.
// No CPU fallback: every computing entry point launches CUDA kernels on the context's device/stream.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <set>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/collider_b200.h"
#include "kernels.cuh"
#include "single.cuh"
#include "sort.cuh"
#include "tile.cuh"
#include "drain.cuh"
#include "migrate.cuh"

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

enum { ST_PRE = 0, ST_STAGE_A, ST_EXCH, ST_SCAN, ST_SCATTER, ST_CAND, ST_SOLVE, ST_TAIL, ST_COUNT };
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
    DevBuf<HbRec> rec;       // one device: rec[row]; sharded: the visitor table (own records, then received ones)
    DevBuf<int4> bink;       // (sx, sy, span, kgb) per record
    // table order (kernels.cuh "Table order"): label[row] = HbId rank (sharded: the id), ord[row] = index in the caller's
    // arrays (one device: the same buffer as label), row_of[label] = row (one device only)
    DevBuf<uint32_t> label, ord, row_of;
    DevBuf<double> col_alt[11];
    DevBuf<uint8_t> kind_alt, reit_alt;
    DevBuf<uint32_t> group_alt, mask_alt, label_alt, ord_alt;
    DevBuf<uint32_t> sort_key, src_row;
    DevBuf<double> tmp_col;  // download_state staging
    bool need_resort = false;
    int resort_period = 16;  // bulk steps between re-sorts (0 = only after upload); CB200_RESORT_PERIOD overrides
    uint32_t steps_since_resort = 0;
    uint64_t resorts = 0;
    uint8_t* collider_dirty = nullptr;        // incremental API (collider.inl): dirty-row flags, emptied by a bulk step
    uint32_t* collider_dirty_count = nullptr;
    unsigned long long* d_sort_cursor = nullptr;  // entry cursor of the re-sort's counting sort
    bool stage_timing = true;  // CUDA events between the stages (diagnostic); off: only around the whole step
    // CUDA graph replay of the step (one launch instead of 6-7); CB200_GRAPH=0 disables
    StepDyn* d_dyn = nullptr;
    StepDyn* h_dyn = nullptr;  // pinned
    bool use_graph = true;
    struct GraphSlot {
        uint64_t sig = 0;
        cudaGraphExec_t exec = nullptr;
    };
    static constexpr int kGraphSlots = 32;  // column-buffer set x step parity (sharded) x dense path on / off x (window drain) overlap-list buffer
    GraphSlot graphs[kGraphSlots];
    int graph_next = 0;
    bool capturing = false;
    // add mode of the step (collider.inl add_hitboxes): see SolveArgs
    bool add_mode = false, last_add_mode = false;
    DevBuf<uint2> imm_pairs;
    unsigned long long* d_n_imm = nullptr;
    uint32_t graph_kernels[kGraphSlots] = {};  // kernels inside each captured graph (for the launch counter)
    uint64_t graph_launches = 0, graph_captures = 0;
    bool solve_two_phase = false;  // CB200_SOLVE=2: k_solve2, filter + dense solve (measured slower: 89 vs 79 us at config 3; kept for A/B and as a second parity check)
    bool ld256 = false;        // CB200_LD256=1: k_solve gathers records with 256-bit loads (measured: 86 vs 84 us, kept for A/B)
    bool stage_a_tma = false;  // CB200_STAGE_A=tma: input columns staged through shared memory by bulk async copies (measured slower, kept for A/B)
    // sharding
    bool sharded = false;
    ShardGeom shard{};
    DevBuf<HbRec> send;      // [slab]: the outgoing halo slab, record 0 is its header
    DevBuf<uint32_t> send_count;
    // owner migration: the own records of the visitor table have room for own_cap rows (the receive area starts there, so that
    // rows can arrive without moving what the peers have mapped)
    uint32_t own_cap = 0, own_cap_request = 0;
    DevBuf<unsigned char> mig_stage;
    DevBuf<uint32_t> mig_u32;
    uint64_t rows_migrated_out = 0, rows_migrated_in = 0;
    uint32_t slab = 0;       // records per slab, header included
    MinKey* d_key = nullptr; // local minimum of the last step (device), input of the all-gather min
    bool own_stream = true;
    DevBuf<uint32_t> vmap;   // gid -> record index (direct-addressed over the id space)
    uint64_t id_space = 0;
    // direct exchange over peer memory (kernels.cuh "Direct halo exchange")
    bool p2p = false;
    bool p2p_same_process = false;  // a peer context lives in this process (tests): no graph capture between steps that wait for each other
    PeerTable peers{};
    Mailbox* d_mail = nullptr;
    void* ipc_opened[2 * kMaxWorld] = {};
    int n_ipc_opened = 0;
    uint64_t p2p_step = 0;
    int p2p_parity = 0;
    GlobalKey* h_gkey = nullptr;  // pinned, mapped
    GlobalKey* d_gkey = nullptr;
    unsigned long long peer_timeout_ns = 20ull * 1000000000ull;  // bound of every wait for a peer (CB200_PEER_TIMEOUT_MS)
    int shard_phase = 0;     // 0 idle, 1 after shard_begin, 2 after shard_finish (result pending)
    double shard_T = 0;

    GridGeom geom{0, 0};
    bool geom_forced = false;
    double world_cells = 0;  // cells of the uploaded extent (choose_geom): the density the tile size is chosen from
    // ---- tile path (tile.cuh): binning / enumeration / solve per block of cells in shared memory ----
    bool tile_enabled = false;    // CB200_TILE=1: single-device bulk steps take the tile path
    bool tile_ok = false;         // the row-range / home tables match the current row order (written at the re-sort)
    bool tile_suspended = false;  // a tile step exceeded a capacity: global path until the next re-sort
    bool step_is_tile = false, last_step_tile = false;
    bool global_grid_valid = false;  // the slot table / entry array / records describe the last step (always after a global step)
    TileGeom tgeom{};
    uint32_t tile_target = 48;    // home rows per tile the tile size is chosen for (CB200_TILE_TARGET)
    uint32_t tile_cap_rec = 96, tile_cap_ent = 384, tile_vcap = 0;  // per warp (CB200_TILE_CAP)
    uint32_t per_sm_t = CB200_LB_TILE, grid_t = 0;
    DevBuf<uint2> surv;           // survivors of the tile kernel's filter: (source hi, source lo) (tile.cuh SurvList)
    size_t tile_smem = 0;
    DevBuf<uint2> tile_rows;
    DevBuf<uint32_t> home, vis_count;
    DevBuf<HbRec> vis;
    uint64_t tile_steps = 0, tile_fallbacks = 0;
    DevBuf<uint32_t> slots;
    DevBuf<uint32_t> chunk_base;  // start of each scan chunk's runs in the entry array
    DevBuf<CellEntry> entries;
    DevBuf<WorkItem> work;         // (record, record, slot) work items: kWorkLists equal segments (kernels.cuh WorkLists)
    DevBuf<uint32_t> work_count;   // [kWorkLists + 2]: item counts of the work lists, the dense-run descriptor count, k_solve_dense's cursor
    DevBuf<uint2> dense;           // dense-run descriptors (k_enum_pairs -> k_solve_dense)
    uint32_t grid_d = 0;           // blocks of k_solve_dense
    // Dense slot runs: from rank dense_rank on, the entries of a run go to k_solve_dense instead of becoming work items
    // (kDenseOff: none do).  Both paths are exact; which one a step uses is decided from the PREVIOUS step's counters
    // (choose_dense): in a sparse scene the dense kernel would be an empty launch in front of k_solve (measured 6 us alone,
    // more inside the step), in a dense one it is ~2x faster than work items.  CB200_DENSE_RANK pins the choice (0: off).
    uint32_t dense_rank = kDenseOff;
    uint32_t last_dense_rank = kDenseOff;  // what the last completed step ran with (cb200_fetch_candidate_pairs)
    uint32_t dense_rank_cfg = kDenseRank;
    bool dense_auto = true;
    double dense_auto_ratio = 2.0;     // dense when the last step solved more than this many candidate pairs per cell entry

    DevBuf<uint2> ov_pairs;
    DevBuf<unsigned long long> ov_table;  // buckets of four keys (kernels.cuh OverlapSet)
    DevBuf<uint32_t> ov_bits;             // one bit per gid: has at least one overlap
    DevBuf<uint32_t> ov_pair_bits;        // one bit per hash(pair), >= 16 bits per pair
    uint32_t ov_pair_mask = 0;            // its words - 1
    uint32_t n_ov = 0, ov_bucket_mask = 0;
    DevBuf<uint64_t> ov_ida, ov_idb;
    // window drain (drain.cuh): the overlap set lives on the device from step to step
    bool ov_track = false, ov_ready = false;
    OvDyn* d_ov = nullptr;
    DevBuf<uint2> ov_pairs_new;   // compaction target / sort scratch
    DevBuf<uint32_t> ov_sort_hist;
    bool ov_sort = false;         // CB200_OV_SORT=1: keep the overlapping-pair list in the row order of its first hitbox (measured: 1 % on the dense solve stage)
    uint32_t ov_cap_pairs = 0, ov_n_keys = 0, ov_bit_words = 0;
    uint32_t ov_list_len = 0;     // host mirror of OvDyn.n_pairs (the list also holds the pairs that separated since the last rebuild)
    uint64_t drain_reiterates = 0;  // Reiterate events that fell inside a drained window (the drain is exact only without them)

    DevBuf<PairEvent> events;
    DevBuf<MinKey> partial;
    uint32_t grid_a = 0, grid_p = 0, grid_s = 0;        // blocks of stage A, of the enumeration, of the solve stage
    uint32_t per_sm_a = 8, per_sm_p = 8, per_sm_s = 4;  // ... per SM (CB200_PER_SM_A / _P / _S); solve: one wave (measured 84 -> 79 us)
    uint32_t per_sm_d = CB200_LB_DENSE2;                  // k_solve_dense: one wave at its launch bounds (CB200_PER_SM_D)
    bool dense_lockstep = false;                          // CB200_DENSE_KERNEL=1: the first dense kernel (A/B)
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
    DevBuf<uint32_t> bs_tab[5];  // bucket sort, one per recursion depth: counts / offsets, cursors, oversized buckets, range
    DevBuf<EventOut> ev_out;
    Key128* sorted_keys = nullptr;  // where the last sort left the sorted keys (keys_a or keys_b)
    bool ev_out_valid = false;      // ev_out holds the 32-byte records of the sorted keys
    int prio_low = 0;
    cudaEvent_t ev_keys_done = nullptr;
    bool ev_sorted = false;
    // pipelined I/O: compact events leave on their own copy stream (double-buffered), staged velocities arrive on another
    DevBuf<Event16> ev16[2];
    int ev16_parity = 0;
    cudaStream_t d2h_stream = nullptr, h2d_stream = nullptr;
    cudaEvent_t ev_sorted_done = nullptr, ev_staged_done[2] = {nullptr, nullptr};
    bool fetch_pending = false, fetch_oneshot_pending = false;
    Event16* fetch_dst = nullptr;
    uint64_t fetch_take = 0;
    uint32_t fetch_nt = 0, fetch_np = 0;
    uint64_t fetch_ns = 0;
    DevBuf<double> nvel_staged[2][5];  // two staged sets: one can be uploaded while the step consumes the other
    bool staged_has[2][5] = {};
    int staged_count = 0, staged_head = 0;  // FIFO over the two sets
    bool sort_oneshot = true;        // CB200_SORT=slow: recursive bucket sort / radix sort only (A/B and tests)
    uint32_t* h_sort_big = nullptr;  // pinned, mapped: k_bs_scan8's oversized-bucket report [1 + 2 * kBsMaxBig]
    uint32_t* d_sort_big = nullptr;  // its device alias
    // pipelined fetch: the one-shot sort as two replayed graphs (keys on the context's stream, sort + decode on the copy stream)
    DevBuf<SortDyn> sort_dyn;
    struct FetchGraph {
        uint64_t sig = 0;
        cudaGraphExec_t keys = nullptr, sort = nullptr;
    };
    static constexpr int kFetchGraphs = 8;  // column-buffer set x destination buffer (+ re-captures after a buffer grew)
    FetchGraph fetch_graphs[kFetchGraphs];
    int fetch_graph_next = 0;
    bool fetch_graph_enabled = true;
    bool solve_prefetch = false;  // CB200_SOLVE_PREFETCH=1
    // the copy-stream half of a pipelined fetch (sort graph + download) is launched AFTER the next step has been queued, so
    // that the host's driver calls for it are off the step-to-step critical path (flush_deferred_fetch)
    cudaGraphExec_t fetch_deferred_sort = nullptr;
    const Event16* fetch_deferred_src = nullptr;
    uint64_t sort_fallbacks = 0;
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
    if (f & kFPeerLost) add("a peer rank did not deliver its halo records or key (aborted or timed out)");
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
    c->world_cells = 0;
    if (c->geom_forced) return;
    double minx = INFINITY, maxx = -INFINITY, miny = INFINITY, maxy = -INFINITY;
    for (uint64_t i = 0; i < v->n; ++i) {
        minx = std::min(minx, v->pos_x[i]); maxx = std::max(maxx, v->pos_x[i]);
        miny = std::min(miny, v->pos_y[i]); maxy = std::max(maxy, v->pos_y[i]);
    }
    if (!(maxx >= minx)) minx = maxx = miny = maxy = 0;
    const double cx = std::min((maxx - minx) / c->cell_width + 4.0, 1e9), cy = std::min((maxy - miny) / c->cell_width + 4.0, 1e9);
    c->world_cells = cx * cy;
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
    {
        // the step runs at the highest stream priority, the event sort / download stream (ensure_io_streams) at the lowest: when
        // both have blocks pending, the step's kernels — the critical path, and in sharded mode what the peers wait for — go first
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) least = greatest = 0;
        if (getenv("CB200_NO_PRIORITY")) least = greatest = 0;
        c->prio_low = least;
        if ((e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, greatest)) != cudaSuccess) return bail("cudaStreamCreate", e);
    }
    if ((e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return bail("cudaDeviceGetAttribute", e);
    if ((e = cudaMalloc(&c->d_ctr, sizeof(Counters))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&c->d_key, sizeof(MinKey))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&c->d_sort_cursor, 8)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&c->d_dyn, sizeof(StepDyn))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaHostAlloc(&c->h_dyn, sizeof(StepDyn), cudaHostAllocDefault)) != cudaSuccess) return bail("cudaHostAlloc", e);
    if (const char* g = getenv("CB200_GRAPH")) c->use_graph = g[0] != '0';
    if (const char* rp = getenv("CB200_RESORT_PERIOD")) c->resort_period = std::max(0, atoi(rp));
    if (const char* pt = getenv("CB200_PEER_TIMEOUT_MS")) c->peer_timeout_ns = (unsigned long long)std::max(1, atoi(pt)) * 1000000ull;
    if (const char* sa = getenv("CB200_STAGE_A")) c->stage_a_tma = strcmp(sa, "tma") == 0;
    if (const char* w = getenv("CB200_LD256")) c->ld256 = w[0] != '0';
    if (const char* w = getenv("CB200_SOLVE")) c->solve_two_phase = w[0] == '2';
    if (const char* v = getenv("CB200_DENSE_RANK")) {
        c->dense_auto = false;
        c->dense_rank = atoi(v) <= 0 ? kDenseOff : (uint32_t)atoi(v);
    }
    if (const char* v = getenv("CB200_DENSE_AUTO_RANK")) c->dense_rank_cfg = (uint32_t)std::max(1, atoi(v));
    if (const char* v = getenv("CB200_DENSE_AUTO_RATIO")) c->dense_auto_ratio = atof(v);
    if (const char* v = getenv("CB200_PER_SM_A")) c->per_sm_a = (uint32_t)std::max(1, std::min(32, atoi(v)));
    if (const char* v = getenv("CB200_PER_SM_P")) c->per_sm_p = (uint32_t)std::max(1, std::min(32, atoi(v)));
    if (const char* v = getenv("CB200_PER_SM_S")) c->per_sm_s = (uint32_t)std::max(1, std::min(32, atoi(v)));
    if (const char* v = getenv("CB200_DENSE_KERNEL")) c->dense_lockstep = v[0] == '1';
    if (c->dense_lockstep) c->per_sm_d = CB200_LB_DENSE;
    if (const char* v = getenv("CB200_PER_SM_D")) c->per_sm_d = (uint32_t)std::max(1, std::min(32, atoi(v)));
    if ((e = cudaFuncSetAttribute(k_stage_a_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(StageATile)))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_stage_a_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(StageATile)))) != cudaSuccess)
        return bail("cudaFuncSetAttribute", e);
    if ((e = cudaHostAlloc(&c->h_res, sizeof(StepResultDev), cudaHostAllocMapped)) != cudaSuccess) return bail("cudaHostAlloc", e);
    if ((e = cudaHostGetDevicePointer((void**)&c->d_res, c->h_res, 0)) != cudaSuccess) return bail("cudaHostGetDevicePointer", e);
    if ((e = cudaHostAlloc(&c->h_sort_big, (1 + 2 * kBsMaxBig) * sizeof(uint32_t), cudaHostAllocMapped)) != cudaSuccess) return bail("cudaHostAlloc", e);
    if ((e = cudaHostGetDevicePointer((void**)&c->d_sort_big, c->h_sort_big, 0)) != cudaSuccess) return bail("cudaHostGetDevicePointer", e);
    if (const char* v = getenv("CB200_FETCH_GRAPH")) c->fetch_graph_enabled = v[0] != '0';
    if (const char* v = getenv("CB200_SOLVE_PREFETCH")) c->solve_prefetch = v[0] != '0';
    if (const char* so = getenv("CB200_SORT")) c->sort_oneshot = strcmp(so, "slow") != 0;
    if (const char* v = getenv("CB200_TILE")) c->tile_enabled = v[0] != '0';
    if (const char* v = getenv("CB200_OV_SORT")) c->ov_sort = v[0] != '0';
    if (const char* v = getenv("CB200_TILE_TARGET")) c->tile_target = (uint32_t)std::max(16, std::min(2048, atoi(v)));
    if (const char* v = getenv("CB200_TILE_CAP")) {
        c->tile_cap_rec = (uint32_t)std::max(16, std::min(512, atoi(v)));
        c->tile_cap_ent = 4 * c->tile_cap_rec;
    }
    if (const char* v = getenv("CB200_PER_SM_T")) c->per_sm_t = (uint32_t)std::max(1, std::min(16, atoi(v)));
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
    c->kind.release(); c->reit.release(); c->group.release(); c->mask.release(); c->ids.release(); c->rec.release(); c->bink.release();
    c->label.release(); c->ord.release(); c->row_of.release(); c->send.release(); c->send_count.release(); c->vmap.release(); c->mig_stage.release(); c->mig_u32.release();
    for (auto& b : c->col_alt) b.release();
    c->kind_alt.release(); c->reit_alt.release(); c->group_alt.release(); c->mask_alt.release(); c->label_alt.release(); c->ord_alt.release();
    c->sort_key.release(); c->src_row.release(); c->tmp_col.release();
    if (c->d_sort_cursor) cudaFree(c->d_sort_cursor);
    if (c->d_dyn) cudaFree(c->d_dyn);
    if (c->h_dyn) cudaFreeHost(c->h_dyn);
    for (auto& g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    c->slots.release(); c->chunk_base.release(); c->entries.release(); c->work.release(); c->work_count.release(); c->dense.release();
    c->ov_pairs_new.release(); c->ov_sort_hist.release(); if (c->d_ov) cudaFree(c->d_ov);
    c->ov_pairs.release(); c->ov_table.release(); c->ov_bits.release(); c->ov_pair_bits.release(); c->ov_ida.release(); c->ov_idb.release();
    c->events.release(); c->partial.release();
    c->tile_rows.release(); c->home.release(); c->vis_count.release(); c->vis.release(); c->surv.release();
    for (auto& b : c->ev16) b.release();
    for (auto& set : c->nvel_staged)
        for (auto& b : set) b.release();
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    for (cudaEvent_t e : {c->ev_sorted_done, c->ev_staged_done[0], c->ev_staged_done[1], c->ev_keys_done})
        if (e) cudaEventDestroy(e);
    c->keys_a.release(); c->keys_b.release(); c->rs_hist.release(); for (auto& b : c->bs_tab) b.release(); c->imm_pairs.release(); if (c->d_n_imm) cudaFree(c->d_n_imm); c->ev_out.release();
    if (c->d_ctr) cudaFree(c->d_ctr);
    if (c->d_key) cudaFree(c->d_key);
    for (int i = 0; i < c->n_ipc_opened; ++i) cudaIpcCloseMemHandle(c->ipc_opened[i]);
    if (c->d_mail) cudaFree(c->d_mail);
    if (c->h_gkey) cudaFreeHost(c->h_gkey);
    if (c->h_res) cudaFreeHost(c->h_res);
    if (c->h_sort_big) cudaFreeHost(c->h_sort_big);
    for (auto& g : c->fetch_graphs) {
        if (g.keys) cudaGraphExecDestroy(g.keys);
        if (g.sort) cudaGraphExecDestroy(g.sort);
    }
    c->sort_dyn.release();
    if (c->pinned_stage) cudaFreeHost(c->pinned_stage);
    for (auto& ev : c->ev_t)
        if (ev) cudaEventDestroy(ev);
    if (c->stream && c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int cb200_set_grid(cb200_ctx* ctx, int log2_w, int log2_h) {
    if (!ctx) return CB200_E_ARG;
    if (log2_w == 0 && log2_h == 0) {
        ctx->geom_forced = false;
        return CB200_OK;
    }
    if (log2_w < 1 || log2_h < 1 || log2_w > 20 || log2_h > 20 || log2_w + log2_h < 12 || log2_w + log2_h > 26)
        return fail(ctx, CB200_E_ARG, "grid period out of range (need 12 <= log2_w + log2_h <= 26)");
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

int cb200_set_resort_period(cb200_ctx* ctx, int steps) {
    if (!ctx || steps < 0) return CB200_E_ARG;
    ctx->resort_period = steps;
    return CB200_OK;
}
uint64_t cb200_resort_count(const cb200_ctx* ctx) { return ctx ? ctx->resorts : 0; }
uint64_t cb200_event_sort_fallbacks(const cb200_ctx* ctx) { return ctx ? ctx->sort_fallbacks : 0; }

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
    const size_t n = (size_t)v->n;
    if (ctx->sharded) {
        // room for rows that migrate in later (cb200_shard_add_rows): 1/4 more than the residents of now, or what the host asked for
        const size_t want = ctx->own_cap_request ? (size_t)ctx->own_cap_request : n + std::max<size_t>(n / 4, 4096);
        if (want < n || want >= 0x7fffffffull) return fail(ctx, CB200_E_ARG, "row capacity (cb200_shard_set_row_capacity) smaller than the uploaded table");
        ctx->own_cap = (uint32_t)want;
    }
    const size_t cap = ctx->sharded ? (size_t)ctx->own_cap : std::max<size_t>(n, 1);
    const double* src[11] = {v->pos_x, v->pos_y, v->dim_x, v->dim_y, v->vel_x, v->vel_y, v->rsz_x, v->rsz_y, v->start_time, v->end_time, v->pub_end_time};
    for (int k = 0; k < 11; ++k) {
        CU(ctx->col[k].reserve(cap));
        if (n) CU(cudaMemcpyAsync(ctx->col[k].p, src[k], n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(ctx->kind.reserve(cap)); CU(ctx->reit.reserve(cap)); CU(ctx->group.reserve(cap)); CU(ctx->mask.reserve(cap));
    CU(ctx->ids.reserve(cap)); CU(ctx->rec.reserve(cap)); CU(ctx->bink.reserve(cap));
    if (n) {
        CU(cudaMemcpyAsync(ctx->kind.p, v->kind, n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->group.p, v->group, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->mask.p, v->interact_mask, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->ids.p, v->id, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->reit.p, 0, n, ctx->stream));
    }
    CU(ctx->label.reserve(cap));
    if (ctx->sharded) {
        // sharded tables: label = id (ids must fit 31 bits); visitor table and halo buffers sized here
        std::vector<uint32_t> g(n);
        for (size_t i = 0; i < n; ++i) {
            if (v->id[i] >= ctx->id_space) return fail(ctx, CB200_E_ARG, "sharded mode requires ids < id_space");
            g[i] = (uint32_t)v->id[i];
        }
        CU(ctx->ord.reserve(cap));
        if (n) CU(cudaMemcpyAsync(ctx->label.p, g.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (n) k_iota_rows<<<((uint32_t)n + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>((uint32_t)n, ctx->ord.p, nullptr);
        CU(cudaStreamSynchronize(ctx->stream));
        // halo slab: header + room for n/64 (>= 4095) records leaving the strip (cb200_shard_set_slab overrides); the
        // visitor table holds the n own records followed by one received slab per rank
        if (ctx->slab == 0) ctx->slab = (uint32_t)std::max<size_t>(n / 64, 4095) + 1;
        // (two receive areas: the direct exchange double-buffers them by step parity)
        CU(ctx->rec.reserve(cap + 2 * (size_t)ctx->slab * ctx->shard.world));
        CU(ctx->bink.reserve(cap + 2 * (size_t)ctx->slab * ctx->shard.world));
        ctx->p2p = false;  // peers must be re-imported: the receive area moved
        CU(ctx->send.reserve((size_t)ctx->slab * ctx->shard.world));  // one slab per destination
        CU(ctx->send_count.reserve(kMaxWorld));
        CU(cudaMemsetAsync(ctx->send.p, 0, (size_t)ctx->slab * ctx->shard.world * sizeof(HbRec), ctx->stream));
        CU(ctx->vmap.reserve(ctx->id_space));
        CU(cudaMemsetAsync(ctx->vmap.p, 0xff, ctx->id_space * 4, ctx->stream));
        CU(cudaMemsetAsync(ctx->rec.p + cap, 0, 2 * (size_t)ctx->slab * ctx->shard.world * sizeof(HbRec), ctx->stream));
    } else {
        CU(ctx->row_of.reserve(cap));
        if (n) k_iota_rows<<<((uint32_t)n + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>((uint32_t)n, ctx->label.p, ctx->row_of.p);
    }
    ctx->need_resort = true;
    ctx->steps_since_resort = 0;
    choose_geom(ctx, v);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->n = (uint32_t)n;
    set_l2_window(ctx);
    ctx->have_state = true;
    ctx->have_step = false;
    ctx->ev_sorted = false;
    ctx->n_ov = 0;
    ctx->ov_bucket_mask = 0;
    return CB200_OK;
}

int cb200_download_state(cb200_ctx* ctx, const cb200_state_out* o) {
    if (!ctx || !o) return CB200_E_ARG;
    if (!ctx->have_state) return fail(ctx, CB200_E_STATE, "no state uploaded");
    if (o->n < ctx->n) return fail(ctx, CB200_E_ARG, "state_out capacity too small");
    CU(cudaSetDevice(ctx->device));
    double* dst[11] = {o->pos_x, o->pos_y, o->dim_x, o->dim_y, o->vel_x, o->vel_y, o->rsz_x, o->rsz_y, o->start_time, o->end_time, o->pub_end_time};
    // rows are in grid-slot order: put every column back into the caller's order on the device, then copy
    const uint32_t n = ctx->n;
    if (n) CU(ctx->tmp_col.reserve(n));
    const uint32_t* ord = ctx->sharded ? ctx->ord.p : ctx->label.p;
    for (int k = 0; k < 11; ++k)
        if (dst[k] && n) {
            k_unpermute<<<(n + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(n, ord, ctx->col[k].p, ctx->tmp_col.p);
            ctx->launches += 1;
            CU(cudaMemcpyAsync(dst[k], ctx->tmp_col.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        }
    CU(cudaStreamSynchronize(ctx->stream));
    return CB200_OK;
}

static int gid_to_id_table(cb200_ctx* ctx, std::vector<uint64_t>& ids);
// Put the overlapping-pair list into the row order of its first hitbox (kernels.cuh "Order of the overlapping-pair list").
static int sort_overlap_pairs(cb200_ctx* ctx, uint32_t n_list) {
    if (!ctx->ov_sort || ctx->sharded || n_list < 65536u || ctx->n < 2) return CB200_OK;
    const int shift = 8;
    const uint32_t n_buckets = (ctx->n >> shift) + 2u;
    CU(ctx->ov_sort_hist.reserve(n_buckets));
    CU(ctx->ov_pairs_new.reserve(std::max<size_t>(n_list, ctx->ov_pairs_new.cap)));
    CU(cudaMemsetAsync(ctx->ov_sort_hist.p, 0, (size_t)n_buckets * 4, ctx->stream));
    const uint32_t g = (uint32_t)ctx->sm_count * 8u, n_labels = (uint32_t)std::min<size_t>(ctx->row_of.cap, 0xffffffffu);
    k_ovsort_count<<<g, kThreads, 0, ctx->stream>>>(ctx->ov_pairs.p, n_list, ctx->row_of.p, n_labels, shift, n_buckets, ctx->ov_sort_hist.p);
    k_ovsort_scan<<<1, 1024, 0, ctx->stream>>>(ctx->ov_sort_hist.p, n_buckets);
    k_ovsort_scatter<<<g, kThreads, 0, ctx->stream>>>(ctx->ov_pairs.p, n_list, ctx->row_of.p, n_labels, shift, n_buckets, ctx->ov_sort_hist.p, ctx->ov_pairs_new.p);
    CU(cudaMemcpyAsync(ctx->ov_pairs.p, ctx->ov_pairs_new.p, (size_t)n_list * sizeof(uint2), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->launches += 3;
    return CB200_OK;
}
static int compact_overlaps(cb200_ctx* ctx);
static inline uint64_t id_of(const std::vector<uint64_t>& ids, uint32_t gid);

// (Re)build the overlap hash table, the per-hitbox bits and the pair Bloom bits from the np (hi, lo) label pairs in ov_pairs.
// With the window drain on, everything is sized with head-room (the set changes on the device from step to step).
static int build_overlap_tables(cb200_ctx* ctx, uint32_t np) {
    // cap_pairs: pairs the hash table and the Bloom bits are sized for
    const uint32_t cap_pairs = ctx->ov_track ? (uint32_t)std::max<uint64_t>({4ull * np, (uint64_t)ctx->n / 8, 4096ull}) : np;
    if (ctx->ov_track) {
        // the list can take every pair the table is sized for plus every pair one step can add (one per pair event)
        if (ctx->events.cap == 0) CU(ctx->events.reserve((size_t)ctx->n * 2 + 1024));
        const size_t cap_list = (size_t)cap_pairs + ctx->events.cap;
        CU(ctx->ov_pairs.reserve(cap_list, true, ctx->stream));
        CU(ctx->ov_pairs_new.reserve(cap_list));
        if (!ctx->d_ov) CU(cudaMalloc(&ctx->d_ov, sizeof(OvDyn)));
        OvDyn init{};
        init.n_pairs = np;
        init.n_live = np;
        CU(cudaMemcpyAsync(ctx->d_ov, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));  // (init is a stack object)
        ctx->ov_list_len = np;
    }
    uint32_t cap = 1024;
    while (cap < 2 * cap_pairs) cap <<= 1;
    CU(ctx->ov_table.reserve(cap));
    const size_t gid_space = ctx->sharded ? (size_t)ctx->id_space : std::max<size_t>(ctx->row_of.cap, ctx->n);
    const size_t n_words = (gid_space + 31) / 32 + 1;
    CU(ctx->ov_bits.reserve(n_words));
    CU(cudaMemsetAsync(ctx->ov_bits.p, 0, n_words * 4, ctx->stream));
    uint32_t pair_words = 1024;
    while (pair_words < cap_pairs / 2u + 1u) pair_words <<= 1;  // >= 16 bits per pair
    CU(ctx->ov_pair_bits.reserve(pair_words));
    CU(cudaMemsetAsync(ctx->ov_pair_bits.p, 0, (size_t)pair_words * 4, ctx->stream));
    CU(cudaMemsetAsync(ctx->ov_table.p, 0xff, (size_t)cap * 8, ctx->stream));
    if (np) {
        k_overlap_insert<<<(np + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(ctx->ov_pairs.p, np, ctx->ov_table.p, cap / 4 - 1, ctx->ov_bits.p, ctx->ov_pair_bits.p, pair_words - 1);
        ctx->launches += 1;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    ctx->n_ov = np;
    ctx->ov_bucket_mask = cap / 4 - 1;
    ctx->ov_pair_mask = pair_words - 1;
    ctx->ov_cap_pairs = cap_pairs;
    ctx->ov_n_keys = cap;
    ctx->ov_bit_words = (uint32_t)n_words;
    ctx->ov_ready = ctx->ov_track;
    return ctx->need_resort ? CB200_OK : sort_overlap_pairs(ctx, np);  // (before the first re-sort the rows are still in upload order)
}

// Window drain housekeeping: drop the pairs that separated since the last rebuild from the list and rebuild the tables.
static int compact_overlaps(cb200_ctx* ctx) {
    if (!(ctx->ov_track && ctx->ov_ready)) return CB200_OK;
    const uint32_t len = ctx->ov_list_len;
    uint32_t n_live = 0;
    if (len) {
        uint32_t* d_n = &ctx->d_ov->n_add;  // (a scratch word: zero between steps)
        CU(cudaMemsetAsync(d_n, 0, 4, ctx->stream));
        k_ov_compact<<<(uint32_t)ctx->sm_count * 4u, kThreads, 0, ctx->stream>>>(ctx->ov_pairs.p, len, ctx->ov_table.p, ctx->ov_bucket_mask, ctx->ov_pairs_new.p, d_n);
        ctx->launches += 1;
        CU(cudaMemcpyAsync(&n_live, d_n, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (n_live) CU(cudaMemcpyAsync(ctx->ov_pairs.p, ctx->ov_pairs_new.p, (size_t)n_live * sizeof(uint2), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return build_overlap_tables(ctx, n_live);
}

int cb200_set_overlaps(cb200_ctx* ctx, const uint64_t* id_a, const uint64_t* id_b, uint64_t n_pairs) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->have_state) return fail(ctx, CB200_E_STATE, "set_overlaps before upload_state");
    if (n_pairs && (!id_a || !id_b)) return fail(ctx, CB200_E_ARG, "null id array");
    if (n_pairs >= 0x40000000ull) return fail(ctx, CB200_E_ARG, "too many overlapping pairs");
    CU(cudaSetDevice(ctx->device));
    ctx->n_ov = 0;
    ctx->ov_bucket_mask = 0;
    ctx->ov_ready = false;
    if (n_pairs == 0) return ctx->ov_track ? build_overlap_tables(ctx, 0) : CB200_OK;
    const uint32_t np = (uint32_t)n_pairs;
    CU(ctx->ov_ida.reserve(np)); CU(ctx->ov_idb.reserve(np)); CU(ctx->ov_pairs.reserve(np));
    CU(cudaMemcpyAsync(ctx->ov_ida.p, id_a, (size_t)np * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->ov_idb.p, id_b, (size_t)np * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(Counters), ctx->stream));
    const unsigned long long big = ~0ull;
    CU(cudaMemcpyAsync(&ctx->d_ctr->err_slot, &big, 8, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t blocks = (np + kThreads - 1) / kThreads;
    k_overlap_map_ids<<<blocks, kThreads, 0, ctx->stream>>>(ctx->sharded ? nullptr : ctx->ids.p, ctx->n, ctx->ov_ida.p, ctx->ov_idb.p, np, ctx->ov_pairs.p, ctx->d_ctr);
    ctx->launches += 1;
    Counters h;
    CU(cudaMemcpyAsync(&h, ctx->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    if (h.err_flags) return fail(ctx, CB200_E_ARG, "set_overlaps: hitbox id not found (pair " + std::to_string(h.err_slot) + ")");
    return build_overlap_tables(ctx, np);
}

int cb200_set_window_drain(cb200_ctx* ctx, int on) {
    if (!ctx) return CB200_E_ARG;
    if (on && ctx->sharded) return fail(ctx, CB200_E_STATE, "the window drain runs on one device");
    CU(cudaSetDevice(ctx->device));
    const bool was = ctx->ov_track;
    ctx->ov_track = on != 0;
    ctx->ov_ready = false;
    if (ctx->ov_track && !was && ctx->have_state) return build_overlap_tables(ctx, ctx->n_ov);  // (the pairs uploaded so far stay)
    return CB200_OK;
}

// The overlap set as the device holds it (with the window drain: as of the end of the last drained window).
int cb200_fetch_overlaps(cb200_ctx* ctx, uint64_t* id_a, uint64_t* id_b, uint64_t cap, uint64_t* n_out) {
    if (!ctx) return CB200_E_ARG;
    if (ctx->sharded) return fail(ctx, CB200_E_STATE, "fetch_overlaps is a one-device view");
    CU(cudaSetDevice(ctx->device));
    if (ctx->ov_track && ctx->ov_ready && ctx->ov_list_len != ctx->n_ov) {
        int rc = compact_overlaps(ctx);  // (the list also holds the pairs that separated since the last rebuild)
        if (rc != CB200_OK) return rc;
    }
    const uint64_t total = ctx->n_ov;
    if (n_out) *n_out = total;
    if (cap == 0 || total == 0) return CB200_OK;
    std::vector<uint2> h(total);
    std::vector<uint64_t> ids;
    int rc = gid_to_id_table(ctx, ids);
    if (rc != CB200_OK) return rc;
    CU(cudaMemcpy(h.data(), ctx->ov_pairs.p, total * sizeof(uint2), cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < total && i < cap; ++i) {
        if (id_a) id_a[i] = id_of(ids, h[i].x);
        if (id_b) id_b[i] = id_of(ids, h[i].y);
    }
    return CB200_OK;
}

int cb200_last_drain(const cb200_ctx* ctx, uint64_t* followups, uint64_t* added, uint64_t* removed, uint64_t* pairs_now, uint64_t* reiterates_in_windows) {
    if (!ctx) return CB200_E_ARG;
    if (followups) *followups = ctx->last.n_followups;
    if (added) *added = ctx->last.n_added;
    if (removed) *removed = ctx->last.n_removed;
    if (pairs_now) *pairs_now = ctx->n_ov;
    if (reiterates_in_windows) *reiterates_in_windows = ctx->drain_reiterates;
    return CB200_OK;
}

// CB200_SYNC_DEBUG=1 (with CB200_GRAPH=0): wait after every stage and name the one that failed.
static int dbg_sync(cb200_ctx* ctx, const char* stage) {
    static const bool on = getenv("CB200_SYNC_DEBUG") != nullptr;
    if (!on) return CB200_OK;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, CB200_E_CUDA, std::string("stage ") + stage + ": " + cudaGetErrorString(e));
    return CB200_OK;
}
#define DBG_SYNC(stage) do { int rc_ = dbg_sync(ctx, stage); if (rc_ != CB200_OK) return rc_; } while (0)

static int stage_vel(cb200_ctx* ctx, const cb200_vel_view* vel, const double** nv) {
    const uint32_t n = ctx->n;
    for (int k = 0; k < 5; ++k) nv[k] = nullptr;
    if (!vel && ctx->staged_count) {
        // velocities staged ahead of time by cb200_stage_velocities (their H2D copy ran beside the previous step): oldest set first
        const int slot = ctx->staged_head;
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_staged_done[slot], 0));
        for (int k = 0; k < 5; ++k)
            if (ctx->staged_has[slot][k]) nv[k] = ctx->nvel_staged[slot][k].p;
        ctx->staged_head ^= 1;
        ctx->staged_count -= 1;
        return CB200_OK;
    }
    if (vel && n) {
        const double* src[5] = {vel->vel_x, vel->vel_y, vel->rsz_x, vel->rsz_y, vel->end_time};
        for (int k = 0; k < 5; ++k)
            if (src[k]) {
                CU(ctx->nvel[k].reserve(n));
                CU(cudaMemcpyAsync(ctx->nvel[k].p, src[k], (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
                nv[k] = ctx->nvel[k].p;
            }
    }
    return CB200_OK;
}

static ShardGeom whole_world() {
    ShardGeom g{};
    g.world = 1;
    g.rank = 0;
    g.col_lo = INT32_MIN;
    g.col_hi = INT32_MAX;
    g.bounds[0] = INT32_MIN;
    g.bounds[1] = INT32_MAX;
    return g;
}

// Slot counts -> run starts (k_scan_slots).  step = the bulk step's / re-bin's scan (totals go to the step counters);
// otherwise the re-sort's.
static DenseList dense_list_of(cb200_ctx* ctx) {
    return DenseList{ctx->dense.p, ctx->work_count.p + kWorkLists, (uint32_t)std::min<size_t>(ctx->dense.cap, 0xffffffffu), ctx->work_count.p + kWorkLists + 1, ctx->dense_rank};
}
static void choose_dense(cb200_ctx* ctx) {
    if (!ctx->dense_auto) return;
    const Counters& c = ctx->last.ctr;
    // hysteresis: a scene hovering at the threshold must not flip every step (each flip is another step graph)
    const double ratio = ctx->dense_rank == kDenseOff ? ctx->dense_auto_ratio : 0.75 * ctx->dense_auto_ratio;
    const bool dense = ctx->have_step && c.n_entries > 0 && (double)c.n_collide > ratio * (double)c.n_entries;
    ctx->dense_rank = dense ? ctx->dense_rank_cfg : kDenseOff;
}
// one descriptor per long run (> rank entries) + one per 32 entries: never more than this.  Sized for the configured rank even
// while the dense path is switched off (choose_dense): switching it on must not reallocate in the middle of a run — a
// cudaFree / cudaMalloc waits for the device, and with a direct peer exchange in flight (k_wait_flags spinning on a context
// whose peer has not been enqueued yet) that wait would never end.
static cudaError_t reserve_dense(cb200_ctx* ctx) {
    const uint32_t rank = ctx->dense_auto ? ctx->dense_rank_cfg : ctx->dense_rank;
    if (rank == kDenseOff) return ctx->dense.reserve(64);
    return ctx->dense.reserve(ctx->entries.cap / ((size_t)rank + 1) + ctx->entries.cap / 32 + 64);
}
static WorkLists work_lists_of(cb200_ctx* ctx) {
    return WorkLists{ctx->work.p, (uint32_t)std::min<size_t>(ctx->work.cap / kWorkLists, 0x7fffffffu), ctx->work_count.p};
}

static int launch_scan(cb200_ctx* ctx, bool step) {
    const uint32_t n_slots = ctx->geom.n_slots();
    if (!step) CU(cudaMemsetAsync(ctx->d_sort_cursor, 0, 8, ctx->stream));
    k_scan_slots<<<n_slots / kScanChunk, kThreads, 0, ctx->stream>>>(ctx->slots.p, step ? &ctx->d_ctr->n_entries : ctx->d_sort_cursor, ctx->chunk_base.p,
                                                                       step ? &ctx->d_ctr->n_cells : nullptr);
    ctx->launches += 1;
    return CB200_OK;
}

// Tile path: the tile size for this table — about tile_target home rows per tile at the scene's mean density — and the
// buffers that depend on it.  Called at every re-sort (the density may have changed with the table).
static int choose_tile_geom(cb200_ctx* ctx) {
    const GridGeom g = ctx->geom;
    const double n_slots = (double)g.n_slots();
    const double cells = ctx->world_cells > 0 ? std::min(n_slots, ctx->world_cells) : n_slots;
    const double per_cell = std::max(1e-9, (double)ctx->n / cells);
    int tj = (int)lround(log2((double)ctx->tile_target / per_cell));
    tj = std::max(4, std::min(std::min(11, g.lw + g.lh), tj));
    if (const char* v = getenv("CB200_TILE_LOG2")) tj = std::max(1, std::min(std::min(11, g.lw + g.lh), atoi(v)));
    ctx->tgeom = make_tile_geom(g, tj);
    {
        // visitors a tile can expect: the hitboxes whose rect (about 1.45 cells wide at these shape sizes) reaches in from the
        // neighbouring tiles, plus the rows that drift out of their home tile between two re-sorts; twice that, plus slack
        const double own = per_cell * (double)(1u << tj);
        const double wx = (double)(1u << ctx->tgeom.bx), wy = (double)(1u << ctx->tgeom.by);
        const double ratio = (1.0 + 0.6 / wx) * (1.0 + 0.6 / wy) - 1.0;
        ctx->tile_vcap = (uint32_t)std::min(4096.0, std::max(32.0, 2.0 * own * ratio + 32.0));
    }
    if (const char* v = getenv("CB200_TILE_VCAP")) ctx->tile_vcap = (uint32_t)std::max(1, atoi(v));
    const uint32_t nt = ctx->tgeom.n_tiles;
    CU(ctx->tile_rows.reserve(nt));
    CU(ctx->home.reserve(std::max<uint32_t>(ctx->n, 1)));
    CU(ctx->vis_count.reserve(2 * (size_t)nt));
    CU(ctx->vis.reserve((size_t)nt * ctx->tile_vcap));
    CU(cudaMemsetAsync(ctx->vis_count.p, 0, 2 * (size_t)nt * 4, ctx->stream));
    ctx->tile_smem = kTileWarps * tile_warp_smem_bytes(ctx->tile_cap_rec, ctx->tile_cap_ent, tj);
    CU(cudaFuncSetAttribute(k_tile_bin<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->tile_smem));
    CU(cudaFuncSetAttribute(k_tile_bin<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->tile_smem));
    if (ctx->surv.cap == 0) CU(ctx->surv.reserve((size_t)ctx->n + 4096));
    ctx->grid_t = (uint32_t)ctx->sm_count * ctx->per_sm_t;
    if (getenv("CB200_VERBOSE"))
        fprintf(stderr, "[cb200] tile path: 2^%d slots per tile (%d x %d cells), %u tiles, %.2f hitboxes per cell, record cap %u per warp, visitor cap %u, %zu B smem per block, %u blocks\n", tj,
                1 << ctx->tgeom.bx, 1 << ctx->tgeom.by, nt, per_cell, ctx->tile_cap_rec, ctx->tile_vcap, ctx->tile_smem, ctx->grid_t);
    return CB200_OK;
}

// Put the table rows into grid-slot order of their centre cells (kernels.cuh "Table order").
static int resort_rows(cb200_ctx* ctx) {
    const uint32_t n = ctx->n;
    ctx->need_resort = false;
    ctx->steps_since_resort = 0;
    ctx->tile_ok = false;
    if (n < 2) return CB200_OK;
    const uint32_t n_slots = ctx->geom.n_slots();
    const uint32_t blocks = (n + kThreads - 1) / kThreads;
    for (int k = 0; k < 11; ++k) CU(ctx->col_alt[k].reserve(ctx->col[k].cap));
    CU(ctx->kind_alt.reserve(ctx->kind.cap)); CU(ctx->reit_alt.reserve(ctx->reit.cap)); CU(ctx->group_alt.reserve(ctx->group.cap));
    CU(ctx->mask_alt.reserve(ctx->mask.cap)); CU(ctx->label_alt.reserve(ctx->label.cap));
    if (ctx->sharded) CU(ctx->ord_alt.reserve(ctx->ord.cap));
    CU(ctx->sort_key.reserve(n)); CU(ctx->src_row.reserve(n));
    CU(cudaMemsetAsync(ctx->slots.p, 0, (size_t)n_slots * 4, ctx->stream));
    k_resort_count<<<blocks, kThreads, 0, ctx->stream>>>(n, ctx->col[0].p, ctx->col[1].p, ctx->cell_width, ctx->geom, ctx->slots.p, ctx->sort_key.p);
    int rc = launch_scan(ctx, false);
    if (rc != CB200_OK) return rc;
    k_resort_rank<<<blocks, kThreads, 0, ctx->stream>>>(n, ctx->sort_key.p, ctx->slots.p, ctx->src_row.p);
    if (ctx->tile_enabled && !ctx->sharded && n_slots >= 16) {
        // tile path: the row range of every tile and the home tile of every (new) row
        if ((rc = choose_tile_geom(ctx)) != CB200_OK) return rc;
        k_tile_rows<<<(ctx->tgeom.n_tiles + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(ctx->tgeom.n_tiles, ctx->tgeom.tj, ctx->slots.p, ctx->chunk_base.p, ctx->tile_rows.p);
        k_tile_home<<<blocks, kThreads, 0, ctx->stream>>>(n, ctx->tgeom.tj, ctx->src_row.p, ctx->sort_key.p, ctx->home.p);
        ctx->launches += 2;
        ctx->tile_ok = true;
        ctx->tile_suspended = false;
    }
    ResortCols rc_{};
    for (int k = 0; k < 11; ++k) {
        rc_.in[k] = ctx->col[k].p;
        rc_.out[k] = ctx->col_alt[k].p;
    }
    rc_.kind_in = ctx->kind.p; rc_.kind_out = ctx->kind_alt.p;
    rc_.reit_in = ctx->reit.p; rc_.reit_out = ctx->reit_alt.p;
    rc_.group_in = ctx->group.p; rc_.group_out = ctx->group_alt.p;
    rc_.mask_in = ctx->mask.p; rc_.mask_out = ctx->mask_alt.p;
    rc_.label_in = ctx->label.p; rc_.label_out = ctx->label_alt.p;
    rc_.ord_in = ctx->sharded ? ctx->ord.p : nullptr; rc_.ord_out = ctx->sharded ? ctx->ord_alt.p : nullptr;
    rc_.row_of = ctx->sharded ? nullptr : ctx->row_of.p;
    k_resort_gather<<<blocks, kThreads, 0, ctx->stream>>>(n, ctx->src_row.p, rc_);
    ctx->launches += 3;
    for (int k = 0; k < 11; ++k) std::swap(ctx->col[k], ctx->col_alt[k]);
    std::swap(ctx->kind, ctx->kind_alt); std::swap(ctx->reit, ctx->reit_alt); std::swap(ctx->group, ctx->group_alt);
    std::swap(ctx->mask, ctx->mask_alt); std::swap(ctx->label, ctx->label_alt);
    if (ctx->sharded) std::swap(ctx->ord, ctx->ord_alt);
    ctx->resorts += 1;
    {
        const uint32_t n_list = (ctx->ov_track && ctx->ov_ready) ? ctx->ov_list_len : ctx->n_ov;
        int rc2 = sort_overlap_pairs(ctx, n_list);  // (the rows moved: the pair list follows)
        if (rc2 != CB200_OK) return rc2;
    }
    return CB200_OK;
}

// Every device allocation a step (and a re-sort) needs.  cudaMalloc synchronises the device, so contexts that wait for each
// other inside their kernels (direct exchange) must have done this before the first step.
static int reserve_step_buffers(cb200_ctx* ctx) {
    choose_dense(ctx);
    const uint32_t n = ctx->n;
    const uint32_t n_slots = ctx->geom.n_slots();
    CU(ctx->slots.reserve(n_slots));
    CU(ctx->chunk_base.reserve(n_slots / kScanChunk));
    if (ctx->entries.cap == 0) CU(ctx->entries.reserve((size_t)n * 4 + 1024));
    if (ctx->work.cap == 0) CU(ctx->work.reserve(std::max<size_t>((size_t)n * 2 + 1024, (size_t)kWorkLists * 2048)));
    CU(ctx->work_count.reserve(kWorkLists + 2));
    CU(reserve_dense(ctx));
    if (ctx->events.cap == 0) CU(ctx->events.reserve((size_t)n * 2 + 1024));
    CU(ctx->partial.reserve((size_t)ctx->sm_count * (32 + 32 + 32 + 8)));  // worst case of the CB200_PER_SM_* overrides
    const uint32_t row_cap = ctx->sharded ? std::max(n, ctx->own_cap) : n;  // (rows may migrate in: cb200_shard_add_rows)
    for (auto& b : ctx->nvel) CU(b.reserve(std::max<uint32_t>(row_cap, 1)));  // (stage_vel must not allocate while a peer waits in a kernel)
    if (row_cap >= 2) {
        for (int k = 0; k < 11; ++k) CU(ctx->col_alt[k].reserve(ctx->col[k].cap));
        CU(ctx->kind_alt.reserve(ctx->kind.cap)); CU(ctx->reit_alt.reserve(ctx->reit.cap)); CU(ctx->group_alt.reserve(ctx->group.cap));
        CU(ctx->mask_alt.reserve(ctx->mask.cap)); CU(ctx->label_alt.reserve(ctx->label.cap));
        if (ctx->sharded) CU(ctx->ord_alt.reserve(ctx->ord.cap));
        CU(ctx->sort_key.reserve(row_cap)); CU(ctx->src_row.reserve(row_cap));
    }
    return CB200_OK;
}

static int prepare_buffers(cb200_ctx* ctx) {
    choose_dense(ctx);
    const uint32_t n = ctx->n;
    const uint32_t n_slots = ctx->geom.n_slots();
    CU(ctx->slots.reserve(n_slots));
    CU(ctx->chunk_base.reserve(n_slots / kScanChunk));
    if (ctx->entries.cap == 0) CU(ctx->entries.reserve((size_t)n * 4 + 1024));
    if (ctx->work.cap == 0) CU(ctx->work.reserve(std::max<size_t>((size_t)n * 2 + 1024, (size_t)kWorkLists * 2048)));
    CU(ctx->work_count.reserve(kWorkLists + 2));
    CU(reserve_dense(ctx));
    if (ctx->events.cap == 0) CU(ctx->events.reserve((size_t)n * 2 + 1024));
    ctx->grid_a = std::max<uint32_t>(1, std::min<uint32_t>((n + kThreads - 1) / kThreads, ctx->sm_count * ctx->per_sm_a));
    ctx->grid_p = ctx->sm_count * ctx->per_sm_p;
    ctx->grid_s = ctx->sm_count * ctx->per_sm_s;
    ctx->grid_d = ctx->sm_count * ctx->per_sm_d;
    CU(ctx->partial.reserve((size_t)ctx->sm_count * 12 + ctx->grid_s + std::max(ctx->grid_d, (uint32_t)ctx->sm_count * ctx->per_sm_t) + 8));
    // the rows of the previous step's records move with a re-sort: only between steps, never on a re-run
    if (ctx->need_resort || (ctx->resort_period > 0 && ctx->steps_since_resort >= (uint32_t)ctx->resort_period)) {
        int rc = resort_rows(ctx);
        if (rc != CB200_OK) return rc;
    }
    ctx->steps_since_resort += 1;
    return CB200_OK;
}

// Front half: zero the step tables, stage A.  One device: also counts cells.  Sharded: packs records per rank.
// The step's time and scalar end_time reach the kernels through ctx->d_dyn (set_step_dyn), not through their arguments.
static int set_step_dyn(cb200_ctx* ctx, double time, double end_time) {
    ctx->h_dyn->T = time;
    ctx->h_dyn->end_scalar = end_time;
    ctx->h_dyn->tag = ctx->sharded ? ctx->p2p_step : ctx->tile_steps;  // (tile path: the parity of its visitor counters)
    CU(cudaMemcpyAsync(ctx->d_dyn, ctx->h_dyn, sizeof(StepDyn), cudaMemcpyHostToDevice, ctx->stream));
    return CB200_OK;
}

static int enqueue_front(cb200_ctx* ctx, double time, bool do_rebase, const double* const* nv, double end_time) {
    const uint32_t n = ctx->n;
    const GridGeom g = ctx->geom;
    const uint32_t n_slots = g.n_slots();
    k_step_zero<<<ctx->sm_count * 4, kThreads, 0, ctx->stream>>>(reinterpret_cast<uint4*>(ctx->slots.p), n_slots / 4, ctx->d_ctr,
                                                                  ctx->sharded ? ctx->send_count.p : nullptr, ctx->work_count.p, kWorkLists + 2, ctx->sharded ? n : 0u,
                                                                  ctx->sharded ? (uint32_t)ctx->shard.world : 0u);
    ctx->launches += 1;
    DBG_SYNC("k_step_zero");
    if (ctx->collider_dirty && n) {
        CU(cudaMemsetAsync(ctx->collider_dirty, 0, n, ctx->stream));
        CU(cudaMemsetAsync(ctx->collider_dirty_count, 0, 4, ctx->stream));
    }
    if (ctx->stage_timing) CU(cudaEventRecord(ctx->ev_t[ST_STAGE_A], ctx->stream));
    StepArgs a{};
    a.n = n;
    a.rebase = do_rebase ? 1 : 0;
    a.dyn = ctx->d_dyn;
    a.cell_width = ctx->cell_width;
    a.padding = ctx->padding;
    a.s = cols_of(ctx);
    a.nvx = do_rebase ? nv[0] : nullptr; a.nvy = do_rebase ? nv[1] : nullptr;
    a.nrw = do_rebase ? nv[2] : nullptr; a.nrh = do_rebase ? nv[3] : nullptr;
    a.nend = do_rebase ? nv[4] : ctx->col[10].p;  // re-run: the user end_time is the stored pub_end_time
    a.label = ctx->label.p;
    a.ord = ctx->sharded ? ctx->ord.p : ctx->label.p;
    a.rec = ctx->rec.p;
    a.bink = ctx->bink.p;
    a.dbg = getenv("CB200_DBG_A") ? (uint32_t)atoi(getenv("CB200_DBG_A")) : 0u;
    a.slot_count = ctx->slots.p;
    a.geom = g;
    a.ctr = ctx->d_ctr;
    a.partial = ctx->partial.p;
    a.send = ctx->send.p;
    a.send_count = ctx->send_count.p;
    a.slab = ctx->slab;
    a.vmap = (ctx->sharded && ctx->n_ov) ? ctx->vmap.p : nullptr;
    a.ov_bits = ctx->ov_bits.p;
    if (n) {
        // CB200_STAGE_A=tma: full 256-row tiles through the TMA-staged kernel, the tail through the plain one (default: everything)
        const uint32_t n_tiles = ctx->stage_a_tma ? n / kThreads : 0u;
        const ShardGeom shg = ctx->sharded ? ctx->shard : whole_world();
        uint32_t used = 0;
        if (n_tiles) {
            const uint32_t grid_t = std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count * 4u);
            const size_t smem = 2 * sizeof(StageATile);
            a.partial_base = 0;
            if (ctx->sharded) k_stage_a_tma<true><<<grid_t, kThreads, smem, ctx->stream>>>(a, shg, n_tiles);
            else k_stage_a_tma<false><<<grid_t, kThreads, smem, ctx->stream>>>(a, shg, n_tiles);
            used = grid_t;
            ctx->launches += 1;
        }
        const uint32_t first_row = n_tiles * kThreads;
        if (first_row < n) {
            const uint32_t grid_r = std::max<uint32_t>(1, std::min<uint32_t>((n - first_row + kThreads - 1) / kThreads, ctx->sm_count * 8));
            a.partial_base = used;
            if (ctx->sharded) k_stage_a<true><<<grid_r, kThreads, 0, ctx->stream>>>(a, shg, first_row);
            else k_stage_a<false><<<grid_r, kThreads, 0, ctx->stream>>>(a, shg, first_row);
            used += grid_r;
            ctx->launches += 1;
        }
        ctx->grid_a = used;
    } else {
        ctx->grid_a = 1;
        CU(cudaMemsetAsync(ctx->partial.p, 0xff, sizeof(MinKey) * ctx->grid_a, ctx->stream));
    }
    DBG_SYNC("k_stage_a");
    if (ctx->sharded) {
        k_slab_header<<<1, kMaxWorld, 0, ctx->stream>>>(ctx->send.p, ctx->send_count.p, ctx->slab, ctx->shard.world);
        ctx->launches += 1;
    }
    if (ctx->stage_timing) CU(cudaEventRecord(ctx->ev_t[ST_EXCH], ctx->stream));
    return CB200_OK;
}

// Window drain (drain.cuh) behind the solve stage: chains of the step's pair events up to the step's end time, the overlap
// set updated in place — one launch, which does nothing if the step is going to be re-run.
static int enqueue_drain(cb200_ctx* ctx, bool is_tile) {
    if (!(ctx->ov_track && ctx->ov_ready) || ctx->sharded || ctx->add_mode) return CB200_OK;
    DrainArgs d{};
    d.s = cols_of(ctx);
    d.row_of = ctx->row_of.p;
    d.n_rows = ctx->n;
    d.n_labels = (uint32_t)std::min<size_t>(ctx->row_of.cap, 0xffffffffu);
    d.dyn = ctx->d_dyn;
    d.padding = ctx->padding;
    d.events = ctx->events.p;
    d.cap_events = (uint32_t)std::min<size_t>(ctx->events.cap, 0xffffffffu);
    d.ctr = ctx->d_ctr;
    d.table = ctx->ov_table.p;
    d.bucket_mask = ctx->ov_bucket_mask;
    d.bits = ctx->ov_bits.p;
    d.pair_bits = ctx->ov_pair_bits.p;
    d.pair_mask = ctx->ov_pair_mask;
    d.pairs = ctx->ov_pairs.p;
    d.cap_pairs = (uint32_t)std::min<size_t>(ctx->ov_pairs.cap, 0x7fffffffu);
    d.cap_table_pairs = ctx->ov_cap_pairs;
    d.ov = ctx->d_ov;
    d.result = ctx->d_res;
    d.cap_entries = (uint32_t)std::min<size_t>(ctx->entries.cap, 0xffffffffu);
    d.seg_cap = (uint32_t)std::min<size_t>(ctx->work.cap / kWorkLists, 0x7fffffffu);
    d.cap_surv = (uint32_t)std::min<size_t>(ctx->surv.cap, 0x7fffffffu);
    d.is_tile = is_tile ? 1u : 0u;
    d.work_count = ctx->work_count.p;
    k_drain_chains<<<(uint32_t)ctx->sm_count * 2u, kThreads, 0, ctx->stream>>>(d);
    ctx->launches += 1;
    DBG_SYNC("drain");
    return CB200_OK;
}

// The arguments of the solve-stage kernels (k_solve, k_solve_dense, k_tile_solve) for the current step.
static SolveArgs solve_args_of(cb200_ctx* ctx, const ShardGeom& sh, const VisSpace& vs) {
    const bool use_map = ctx->sharded && ctx->n_ov;
    CandArgs cd{};
    cd.geom = ctx->geom;
    cd.col_lo = sh.col_lo;
    cd.col_hi = sh.col_hi;
    const bool track = ctx->ov_track && ctx->ov_ready;  // (the pair count lives on the device: the table is consulted whenever it exists)
    cd.ov = OverlapSet{reinterpret_cast<const ulonglong2*>(ctx->ov_table.p), ctx->ov_bits.p, ctx->ov_bucket_mask, track ? 1u : ctx->n_ov, ctx->ov_pair_bits.p, ctx->ov_pair_mask};
    SolveArgs sv{};
    sv.cd = cd;
    sv.work = work_lists_of(ctx);
    sv.ov_pairs = ctx->ov_pairs.p;
    sv.n_ov = track ? (uint32_t)std::min<size_t>(ctx->ov_pairs.cap, 0x7fffffffu) : ctx->n_ov;  // (tracked: the capacity; the length is on the device)
    sv.n_ov_dev = track ? &ctx->d_ov->n_pairs : nullptr;
    sv.ov_live_check = track ? 1 : 0;
    sv.rec = ctx->rec.p;
    sv.dyn = ctx->d_dyn;
    sv.padding = ctx->padding;
    sv.events = ctx->events.p;
    sv.cap_events = (uint32_t)std::min<size_t>(ctx->events.cap, 0xffffffffu);
    sv.ctr = ctx->d_ctr;
    sv.partial = ctx->partial.p;
    sv.n_partial_a = ctx->grid_a;
    sv.n_partial_dense = 0;
    sv.result = ctx->d_res;
    sv.local_key = ctx->d_key;
    sv.vmap = ctx->sharded ? (use_map ? ctx->vmap.p : nullptr) : ctx->row_of.p;
    sv.id_space = ctx->sharded ? (uint32_t)ctx->id_space : ctx->n;
    sv.sharded = ctx->sharded ? 1 : 0;
    sv.ids = ctx->sharded ? nullptr : ctx->ids.p;
    sv.dbg = getenv("CB200_DBG_S") ? (uint32_t)atoi(getenv("CB200_DBG_S")) : 0u;
    sv.prefetch = ctx->solve_prefetch ? 1u : 0u;
    sv.add_mode = ctx->add_mode ? 1 : 0;
    sv.imm_pairs = ctx->imm_pairs.p;
    sv.n_imm = ctx->d_n_imm;
    sv.cap_imm = (uint32_t)std::min<size_t>(ctx->imm_pairs.cap, 0xffffffffu);
    sv.vs = vs;
    sv.col_lo = sh.col_lo;
    sv.col_hi = sh.col_hi;
    return sv;
}

// Back half over the visitor table: (index received records,) scan, scatter, candidates, solve + min-reduce + result.
// recount = the own records' cells must be counted again (re-run after a buffer grew).
static int enqueue_back(cb200_ctx* ctx, double time, bool recount, bool bin_only = false) {
    const GridGeom g = ctx->geom;
    const uint32_t n_slots = g.n_slots();
    const ShardGeom sh = ctx->sharded ? ctx->shard : whole_world();
    const uint32_t recv_base = ctx->sharded ? ctx->own_cap : ctx->n;
    const VisSpace vs{ctx->n, recv_base + (ctx->p2p ? (uint32_t)ctx->p2p_parity * (uint32_t)sh.world * ctx->slab : 0u), ctx->slab, sh.world, sh.rank};
    const uint32_t n_vis_max = ctx->sharded ? ctx->n + ctx->slab * (uint32_t)sh.world : ctx->n;
    const uint32_t grid_v = std::max<uint32_t>(1, std::min<uint32_t>((n_vis_max + kThreads - 1) / kThreads, ctx->sm_count * 8));
    const bool use_map = ctx->sharded && ctx->n_ov;
    if (ctx->stage_timing) CU(cudaEventRecord(ctx->ev_t[ST_SCAN], ctx->stream));
    if (ctx->sharded) {
        const uint32_t work = recount ? n_vis_max : ctx->slab * (uint32_t)sh.world;
        const uint32_t grid_i = std::max<uint32_t>(1, std::min<uint32_t>((work + kThreads - 1) / kThreads, ctx->sm_count * 8));
        uint32_t* vmap = use_map ? ctx->vmap.p : nullptr;
        if (recount) k_index_visitors<true><<<grid_i, kThreads, 0, ctx->stream>>>(ctx->rec.p, ctx->bink.p, vs, ctx->d_ctr, ctx->slots.p, g, sh.col_lo, sh.col_hi, vmap, ctx->ov_bits.p);
        else k_index_visitors<false><<<grid_i, kThreads, 0, ctx->stream>>>(ctx->rec.p, ctx->bink.p, vs, ctx->d_ctr, ctx->slots.p, g, sh.col_lo, sh.col_hi, vmap, ctx->ov_bits.p);
        ctx->launches += 1;
    }
    {
        int rc = launch_scan(ctx, true);
        if (rc != CB200_OK) return rc;
    }
    DBG_SYNC("k_scan_slots");
    if (ctx->stage_timing) CU(cudaEventRecord(ctx->ev_t[ST_SCATTER], ctx->stream));
    if (n_vis_max) {
        const uint32_t cap_e = (uint32_t)std::min<size_t>(ctx->entries.cap, 0xffffffffu);
        if (ctx->sharded) k_scatter<true><<<grid_v, kThreads, 0, ctx->stream>>>(ctx->n, ctx->rec.p, ctx->bink.p, vs, ctx->slots.p, ctx->entries.p, cap_e, g, sh.col_lo, sh.col_hi, ctx->d_ctr);
        else k_scatter<false><<<grid_v, kThreads, 0, ctx->stream>>>(ctx->n, ctx->rec.p, ctx->bink.p, vs, ctx->slots.p, ctx->entries.p, cap_e, g, sh.col_lo, sh.col_hi, ctx->d_ctr);
        ctx->launches += 1;
    }
    DBG_SYNC("k_scatter");
    if (ctx->stage_timing) CU(cudaEventRecord(ctx->ev_t[ST_CAND], ctx->stream));
    EnumArgs en{};
    en.entries = ctx->entries.p;
    en.cap_entries = (uint32_t)std::min<size_t>(ctx->entries.cap, 0xffffffffu);
    en.work = work_lists_of(ctx);
    en.dense = dense_list_of(ctx);
    en.ctr = ctx->d_ctr;
    if (bin_only) {
        // (the grid alone: the debug views and the incremental API after a tile step)
        CU(cudaEventRecord(ctx->ev_t[ST_COUNT], ctx->stream));
        return CB200_OK;
    }
    k_enum_pairs<<<ctx->grid_p, kThreads, 0, ctx->stream>>>(en);
    DBG_SYNC("k_enum_pairs");
    if (ctx->stage_timing) CU(cudaEventRecord(ctx->ev_t[ST_SOLVE], ctx->stream));
    SolveArgs sv = solve_args_of(ctx, sh, vs);
    sv.n_partial_dense = ctx->dense_rank == kDenseOff ? 0u : ctx->grid_d;
    if (ctx->dense_rank != kDenseOff) {
        // long slot runs (dense cells) first: their minima and counts are folded in by the last block of the kernel below
        DenseArgs da{};
        da.list = en.dense;
        da.entries = reinterpret_cast<const uint2*>(ctx->entries.p);
        da.cap_entries = en.cap_entries;
        da.partial_at = ctx->grid_a + ctx->grid_s;
        if (ctx->dense_lockstep) k_solve_dense_lockstep<<<ctx->grid_d, kThreads, 0, ctx->stream>>>(sv, da);
        else k_solve_dense<<<ctx->grid_d, kDenseThreads, 0, ctx->stream>>>(sv, da);
        ctx->launches += 1;
        DBG_SYNC("k_solve_dense");
    }
    const bool net = ctx->dense_rank != kDenseOff;  // (k_solve's NET)
    if (ctx->solve_two_phase && !sv.dbg) k_solve2<<<ctx->grid_s, kThreads, 0, ctx->stream>>>(sv);
    else if (ctx->ld256 && net) k_solve<true, true><<<ctx->grid_s, kThreads, 0, ctx->stream>>>(sv);
    else if (ctx->ld256) k_solve<true, false><<<ctx->grid_s, kThreads, 0, ctx->stream>>>(sv);
    else if (net) k_solve<false, true><<<ctx->grid_s, kThreads, 0, ctx->stream>>>(sv);
    else k_solve<false, false><<<ctx->grid_s, kThreads, 0, ctx->stream>>>(sv);
    DBG_SYNC("k_solve");
    if (ctx->stage_timing) CU(cudaEventRecord(ctx->ev_t[ST_TAIL], ctx->stream));
    {
        int rc = enqueue_drain(ctx, false);
        if (rc != CB200_OK) return rc;
    }
    CU(cudaEventRecord(ctx->ev_t[ST_COUNT], ctx->stream));
    ctx->launches += 2;
    return CB200_OK;
}

// The whole step on the tile path (tile.cuh): counters, stage A + visitor slabs, per-tile binning / enumeration / solve in
// shared memory, then the overlapping pairs (separate_time) and the last-block min-reduction through k_solve.
static int enqueue_tile_step(cb200_ctx* ctx, bool do_rebase, const double* const* nv) {
    const uint32_t n = ctx->n;
    k_step_zero<<<1, kThreads, 0, ctx->stream>>>(nullptr, 0u, ctx->d_ctr, nullptr, ctx->work_count.p, kWorkLists + 2, 0u, 0u);
    ctx->launches += 1;
    DBG_SYNC("k_step_zero");
    if (ctx->collider_dirty && n) {
        CU(cudaMemsetAsync(ctx->collider_dirty, 0, n, ctx->stream));
        CU(cudaMemsetAsync(ctx->collider_dirty_count, 0, 4, ctx->stream));
    }
    if (ctx->stage_timing) CU(cudaEventRecord(ctx->ev_t[ST_STAGE_A], ctx->stream));
    StepArgs a{};
    a.n = n;
    a.rebase = do_rebase ? 1 : 0;
    a.dyn = ctx->d_dyn;
    a.cell_width = ctx->cell_width;
    a.padding = ctx->padding;
    a.s = cols_of(ctx);
    a.nvx = do_rebase ? nv[0] : nullptr; a.nvy = do_rebase ? nv[1] : nullptr;
    a.nrw = do_rebase ? nv[2] : nullptr; a.nrh = do_rebase ? nv[3] : nullptr;
    a.nend = do_rebase ? nv[4] : ctx->col[10].p;  // re-run: the user end_time is the stored pub_end_time
    a.label = ctx->label.p;
    a.ord = ctx->label.p;
    a.rec = ctx->rec.p;
    a.bink = ctx->bink.p;
    a.geom = ctx->geom;
    a.ctr = ctx->d_ctr;
    a.partial = ctx->partial.p;
    a.partial_base = 0;
    a.ov_bits = ctx->ov_bits.p;
    TileTables tt{ctx->tile_rows.p, ctx->home.p, ctx->vis_count.p, ctx->vis.p, ctx->tile_vcap, ctx->tgeom.n_tiles};
    TileStageArgs ta{ctx->tgeom, tt, (ctx->n_ov || (ctx->ov_track && ctx->ov_ready)) ? 1u : 0u};
    ctx->grid_a = std::max<uint32_t>(1, std::min<uint32_t>((n + kThreads - 1) / kThreads, ctx->sm_count * ctx->per_sm_a));
    k_stage_a_tile<<<ctx->grid_a, kThreads, 0, ctx->stream>>>(a, ta);
    ctx->launches += 1;
    DBG_SYNC("k_stage_a_tile");
    if (ctx->stage_timing) {
        CU(cudaEventRecord(ctx->ev_t[ST_EXCH], ctx->stream));
        CU(cudaEventRecord(ctx->ev_t[ST_SCAN], ctx->stream));
        CU(cudaEventRecord(ctx->ev_t[ST_SCATTER], ctx->stream));
        CU(cudaEventRecord(ctx->ev_t[ST_CAND], ctx->stream));
    }
    const ShardGeom sh = whole_world();
    const VisSpace vs{n, n, 0u, 1, 0};
    TileSolveArgs ts{};
    ts.sv = solve_args_of(ctx, sh, vs);
    ts.s = a.s;
    ts.label = ctx->label.p;
    ts.bink = ctx->bink.p;
    ts.tg = ctx->tgeom;
    ts.tt = tt;
    ts.cap_rec = ctx->tile_cap_rec;
    ts.cap_ent = ctx->tile_cap_ent;
    ts.surv = SurvList{ctx->surv.p, (uint32_t)std::min<size_t>(ctx->surv.cap, 0x7fffffffu)};
    ts.sv.n_partial_dense = 0;
    k_tile_bin<false><<<ctx->grid_t, kTileThreads, ctx->tile_smem, ctx->stream>>>(ts);
    ctx->launches += 1;
    DBG_SYNC("k_tile_bin");
    if (ctx->stage_timing) CU(cudaEventRecord(ctx->ev_t[ST_SOLVE], ctx->stream));
    // the survivors (collide_time) and the overlapping pairs (separate_time), the min-reduction and the step result.  The tile
    // kernel counted the pairs that pass candidate_geom; this one takes the overlapping ones among them off again
    k_tile_finish<false><<<ctx->grid_s, kThreads, 0, ctx->stream>>>(ts, nullptr, nullptr, 0ull);
    ctx->launches += 1;
    DBG_SYNC("k_tile_finish");
    if (ctx->stage_timing) CU(cudaEventRecord(ctx->ev_t[ST_TAIL], ctx->stream));
    {
        int rc = enqueue_drain(ctx, true);
        if (rc != CB200_OK) return rc;
    }
    CU(cudaEventRecord(ctx->ev_t[ST_COUNT], ctx->stream));
    return CB200_OK;
}

// After a tile step the global grid (slot table, entry array, records) is not built.  The debug views and the incremental
// Collider API read it: rebuild it from the rebased table (the global path's stage A without the rebase, scan, scatter).
static int rebuild_global_grid(cb200_ctx* ctx) {
    if (ctx->global_grid_valid) return CB200_OK;
    const double* nv[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    CU(ctx->slots.reserve(ctx->geom.n_slots()));
    CU(ctx->chunk_base.reserve(ctx->geom.n_slots() / kScanChunk));
    CU(ctx->entries.reserve(std::max<size_t>((size_t)ctx->last.ctr.n_entries + 1024, ctx->entries.cap)));
    int rc = enqueue_front(ctx, ctx->last_T, false, nv, 0.0);
    if (rc == CB200_OK) rc = enqueue_back(ctx, ctx->last_T, false, true);
    if (rc != CB200_OK) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    ctx->global_grid_valid = true;
    return CB200_OK;
}

// Wait for the step, grow any buffer that overflowed.  *again = the step must be re-run.
static int complete_step(cb200_ctx* ctx, double time, bool* again) {
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    ctx->last = *ctx->h_res;
    ctx->last_add_mode = ctx->add_mode;
    ctx->last_dense_rank = ctx->dense_rank;
    ctx->last_step_tile = ctx->step_is_tile;
    ctx->global_grid_valid = !ctx->step_is_tile;
    ctx->last_T = time;
    ctx->have_step = true;
    ctx->ev_sorted = false;
    const Counters& c = ctx->last.ctr;
    *again = false;
    if (ctx->ov_track && ctx->ov_ready && !ctx->sharded && !ctx->add_mode && (ctx->last.drain_flags & kDrainGo)) {
        // the window drain ran: the overlap set now holds the pairs overlapping at the end of the window
        ctx->n_ov = ctx->last.ov_live_now;
        ctx->ov_list_len = ctx->last.ov_pairs_now;
        ctx->drain_reiterates += c.n_solitaire;
        if (ctx->last.drain_flags & (kDrainOverflow | kDrainLongChain))
            return fail(ctx, CB200_E_STATE, ((ctx->last.drain_flags & kDrainOverflow)
                                                 ? std::string("window drain: a buffer overflowed; the overlap set is no longer in step — call cb200_set_overlaps")
                                                 : std::string("window drain: a pair changed state more than 16 times inside one window")) +
                                                " (flags " + std::to_string(ctx->last.drain_flags) + ", list " + std::to_string(ctx->last.ov_pairs_now) + " of " +
                                                std::to_string(ctx->ov_pairs.cap) + ", live " + std::to_string(ctx->last.ov_live_now) + ", added " + std::to_string(ctx->last.n_added) +
                                                ", follow-ups " + std::to_string(ctx->last.n_followups) + ", pair events " + std::to_string(c.n_pair_events) + " of " +
                                                std::to_string(ctx->events.cap) + ")");
        // every pair ever listed holds a table slot: rebuild compactly when the table passes half load, when the dead pairs
        // outnumber the live ones, or when the list could not take another step's worth of new pairs
        const uint32_t dead = ctx->ov_list_len - ctx->n_ov;
        if (ctx->ov_list_len > ctx->ov_cap_pairs || dead > std::max<uint32_t>(ctx->n_ov, 4096u) || (size_t)ctx->ov_list_len + ctx->events.cap > ctx->ov_pairs.cap) {
            int rc = compact_overlaps(ctx);
            if (rc != CB200_OK) return rc;
        }
    }
    if (ctx->step_is_tile) {
        if (c.tile_fail) {
            // a visitor slab or a single cell did not fit: this step, and every step up to the next re-sort, takes the global path
            ctx->tile_suspended = true;
            ctx->tile_fallbacks += 1;
            CU(cudaMemsetAsync(ctx->vis_count.p, 0, 2 * (size_t)ctx->tgeom.n_tiles * 4, ctx->stream));
            if (getenv("CB200_VERBOSE"))
                fprintf(stderr, "[cb200] tile path: capacity exceeded (code %u: 1 visitor slab, 2 records of one cell, 4 entries of one cell, 8 split stack), global path until the next re-sort\n",
                        c.tile_fail);
            *again = true;
            return CB200_OK;
        }
        if (c.n_surv > ctx->surv.cap) {
            CU(ctx->surv.reserve((size_t)c.n_surv + c.n_surv / 4 + 1024));
            *again = true;
        }
        if (c.n_pair_events > ctx->events.cap) {
            CU(ctx->events.reserve((size_t)c.n_pair_events + c.n_pair_events / 4 + 1024));
            *again = true;
        }
        if (*again) CU(cudaMemsetAsync(ctx->vis_count.p, 0, 2 * (size_t)ctx->tgeom.n_tiles * 4, ctx->stream));
        if (*again && ctx->ov_track && ctx->ov_ready) {  // (the drain's list is sized from the event buffer)
            CU(ctx->ov_pairs.reserve((size_t)ctx->ov_cap_pairs + ctx->events.cap, true, ctx->stream));
            CU(ctx->ov_pairs_new.reserve((size_t)ctx->ov_cap_pairs + ctx->events.cap));
        }
        return CB200_OK;
    }
    if (c.overflow_entries || c.n_entries > ctx->entries.cap) {
        CU(ctx->entries.reserve((size_t)c.n_entries + c.n_entries / 4 + 1024));
        CU(reserve_dense(ctx));
        *again = true;
    }
    if (c.work_overflow > ctx->work.cap / kWorkLists) {  // the fullest work sub-list did not fit its segment
        CU(ctx->work.reserve((size_t)kWorkLists * ((size_t)c.work_overflow + c.work_overflow / 4 + 64)));
        *again = true;
    }
    if (c.n_pair_events > ctx->events.cap) {
        CU(ctx->events.reserve((size_t)c.n_pair_events + c.n_pair_events / 4 + 1024));
        *again = true;
    }
    if (*again && ctx->ov_track && ctx->ov_ready) {  // (the drain's list is sized from the event buffer)
        CU(ctx->ov_pairs.reserve((size_t)ctx->ov_cap_pairs + ctx->events.cap, true, ctx->stream));
        CU(ctx->ov_pairs_new.reserve((size_t)ctx->ov_cap_pairs + ctx->events.cap));
    }
    return CB200_OK;
}

static void accumulate_timing(cb200_ctx* ctx, int first_stage, bool with_total) {
    float ms = 0;
    for (int s = first_stage; s < ST_COUNT && ctx->stage_timing; ++s)
        if (cudaEventElapsedTime(&ms, ctx->ev_t[s], ctx->ev_t[s + 1]) == cudaSuccess) {
            ctx->stage_ms_sum[s] += ms;
            if (!with_total) ctx->total_ms_sum += ms;
        }
    if (with_total && cudaEventElapsedTime(&ms, ctx->ev_t[ST_PRE], ctx->ev_t[ST_COUNT]) == cudaSuccess) ctx->total_ms_sum += ms;
}

static int fill_result(cb200_ctx* ctx, cb200_step_result* out) {
    const StepResultDev& r = ctx->last;
    const uint32_t n = ctx->n;
    if (out) {
        memset(out, 0, sizeof(*out));
        out->next_time = r.next_time;
        out->next_event.time = r.next_time;
        if (r.next_tie != kKeyMax) {
            const bool collide = ctx->last_add_mode ? (r.next_tie & 1ull) != 0 : (r.next_tie & kCollideBit) != 0;
            out->next_event.kind = !(r.next_tie & kPairClass) ? CB200_EV_REITERATE : (collide ? CB200_EV_COLLIDE : CB200_EV_SEPARATE);
            out->next_event.a = r.next_id_a;  // (resolved label -> id by the last block of k_solve)
            out->next_event.b = r.next_id_b;
        }
        out->n_hitboxes = n;
        out->n_entries = r.ctr.n_entries;
        out->n_cells = r.ctr.n_cells;
        out->n_collide_solves = r.ctr.n_collide;
        out->n_separate_solves = r.ctr.n_separate;
        out->n_solitaire = r.ctr.n_solitaire;
        out->n_work_items = r.ctr.n_work;
        out->n_events = r.ctr.n_pair_events + r.ctr.n_solitaire;
        out->error_flags = r.ctr.err_flags;
        out->error_slot = 0;
    }
    if (r.ctr.err_flags) {
        uint64_t bad_id = 0;
        if (r.ctr.err_slot < n && !ctx->sharded) CU(cudaMemcpy(&bad_id, ctx->ids.p + r.ctr.err_slot, 8, cudaMemcpyDeviceToHost));
        else bad_id = r.ctr.err_slot;
        if (out) out->error_slot = bad_id;
        return fail(ctx, CB200_E_CONTRACT, flags_message(r.ctr.err_flags) + " (hitbox id " + std::to_string(bad_id) + ")");
    }
    return CB200_OK;
}

// The kernels of one step of a direct-exchange rank: front half, halo push / wait, back half, key push / reduce.
static int enqueue_direct_step(cb200_ctx* ctx, double time, const double* const* nv, double end_time) {
    const int world = ctx->shard.world, rank = ctx->shard.rank;
    int rc = enqueue_front(ctx, time, true, nv, end_time);
    if (rc != CB200_OK) return rc;
    if (world > 1) {
        k_push_halo<<<world - 1, 1024, 0, ctx->stream>>>(ctx->send.p, ctx->peers, world, rank, ctx->slab, ctx->p2p_parity, ctx->d_dyn);
        k_wait_flags<<<1, kMaxWorld, 0, ctx->stream>>>(ctx->d_mail, world, rank, ctx->d_dyn, ctx->d_ctr, ctx->peer_timeout_ns);
        ctx->launches += 2;
    }
    if ((rc = enqueue_back(ctx, time, false)) != CB200_OK) return rc;
    k_push_key<<<1, kMaxWorld, 0, ctx->stream>>>(ctx->d_key, ctx->d_ctr, ctx->work_count.p, (uint32_t)std::min<size_t>(ctx->work.cap / kWorkLists, 0x7fffffffu),
                                                 (uint32_t)std::min<size_t>(ctx->entries.cap, 0xffffffffu), (uint32_t)std::min<size_t>(ctx->events.cap, 0xffffffffu),
                                                 ctx->peers, ctx->d_mail, world, rank, ctx->p2p_parity, ctx->d_dyn);
    k_reduce_keys<<<1, kMaxWorld, 0, ctx->stream>>>(ctx->d_mail, world, rank, ctx->p2p_parity, ctx->d_dyn, ctx->d_gkey, ctx->peer_timeout_ns);
    ctx->launches += 2;
    return CB200_OK;
}

// Everything that is baked into the kernel arguments of a step: if none of it changed since a graph was captured, the graph
// can be replayed (the step's time and scalar end_time travel through d_dyn).
static uint64_t step_signature(cb200_ctx* ctx, const double* const* nv) {
    uint64_t h = 0xcbf29ce484222325ull;
    auto mix = [&h](uint64_t v) {
        for (int i = 0; i < 8; ++i) {
            h ^= (v >> (8 * i)) & 0xff;
            h *= 0x100000001b3ull;
        }
    };
    mix(ctx->n); mix((uint64_t)ctx->geom.lw << 8 | (uint64_t)ctx->geom.lh); mix((ctx->ov_track && ctx->ov_ready) ? 0xffffffffu : ctx->n_ov); mix(ctx->ov_bucket_mask); mix(ctx->ov_pair_mask);
    for (auto& b : ctx->col) mix((uint64_t)(uintptr_t)b.p);
    const void* ptrs[] = {ctx->kind.p, ctx->reit.p, ctx->group.p, ctx->mask.p, ctx->label.p, ctx->row_of.p, ctx->rec.p, ctx->bink.p, ctx->slots.p,
                          ctx->chunk_base.p, ctx->entries.p, ctx->work.p, ctx->work_count.p, ctx->events.p, ctx->partial.p, ctx->ov_pairs.p,
                          ctx->ov_table.p, ctx->ov_bits.p, ctx->ov_pair_bits.p, ctx->ids.p, ctx->collider_dirty, ctx->collider_dirty_count, nv[0], nv[1], nv[2], nv[3], nv[4]};
    for (const void* p : ptrs) mix((uint64_t)(uintptr_t)p);
    mix(ctx->entries.cap); mix(ctx->work.cap); mix(ctx->events.cap); mix((uint64_t)(uintptr_t)ctx->dense.p); mix(ctx->dense.cap); mix(ctx->dense_rank); mix(ctx->dense_lockstep ? 3 : 5); mix(ctx->stage_a_tma ? 1 : 0);
    if (ctx->sharded) {
        mix(0x5aa5); mix(ctx->own_cap); mix(ctx->peer_timeout_ns); mix((uint64_t)ctx->p2p_parity); mix(ctx->slab); mix((uint64_t)ctx->shard.world << 8 | (uint64_t)ctx->shard.rank);
        const void* sp[] = {ctx->ord.p, ctx->send.p, ctx->send_count.p, ctx->vmap.p, ctx->d_mail, ctx->d_key, ctx->d_gkey};
        for (const void* q : sp) mix((uint64_t)(uintptr_t)q);
        for (int d = 0; d < ctx->shard.world; ++d) {
            mix((uint64_t)(uintptr_t)ctx->peers.recv[d]);
            mix((uint64_t)(uintptr_t)ctx->peers.mail[d]);
            mix((uint64_t)(uint32_t)ctx->shard.bounds[d]);
        }
    }
    if (const char* d = getenv("CB200_DBG_A")) mix((uint64_t)atoi(d) + 17);
    if (const char* d = getenv("CB200_DBG_S")) mix((uint64_t)atoi(d) + 31);
    mix(ctx->ld256 ? 257u : 129u);
    mix(ctx->solve_two_phase ? 2u : 1u);
    mix(ctx->solve_prefetch ? 7u : 3u);
    if (ctx->ov_track && ctx->ov_ready) {
        mix(0xd7a1); mix(ctx->ov_cap_pairs); mix(ctx->ov_n_keys); mix(ctx->ov_bit_words); mix(ctx->ov_pairs.cap);
        const void* dp[] = {ctx->d_ov, ctx->surv.p};
        for (const void* q : dp) mix((uint64_t)(uintptr_t)q);
    }
    if (ctx->step_is_tile) {
        mix(0x711e); mix((uint64_t)ctx->tgeom.tj); mix(ctx->tile_vcap); mix(ctx->tile_cap_rec); mix(ctx->tile_cap_ent); mix(ctx->grid_t);
        const void* tp[] = {ctx->tile_rows.p, ctx->home.p, ctx->vis_count.p, ctx->vis.p, ctx->surv.p};
        mix(ctx->surv.cap);
        for (const void* q : tp) mix((uint64_t)(uintptr_t)q);
    }
    return h ? h : 1;
}

// Replays (capturing first, if needed) the whole step — zero, stage A, scan, scatter, enumerate, solve — as ONE graph launch.
static int launch_step_graph(cb200_ctx* ctx, double time, const double* const* nv, double end_time, bool* replayed) {
    *replayed = false;
    const uint64_t sig = step_signature(ctx, nv);
    cb200_ctx::GraphSlot* slot = nullptr;
    for (auto& g : ctx->graphs)
        if (g.exec && g.sig == sig) slot = &g;
    if (!slot) {
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            cudaGetLastError();
            ctx->use_graph = false;  // (e.g. a stream that cannot be captured): plain launches from now on
            return CB200_OK;
        }
        const uint64_t launches_before = ctx->launches;
        ctx->capturing = true;
        int rc;
        if (ctx->sharded) {
            rc = enqueue_direct_step(ctx, time, nv, end_time);
        } else if (ctx->step_is_tile) {
            rc = enqueue_tile_step(ctx, true, nv);
        } else {
            rc = enqueue_front(ctx, time, true, nv, end_time);
            if (rc == CB200_OK) rc = enqueue_back(ctx, time, false);
        }
        ctx->capturing = false;
        const uint32_t n_kernels = (uint32_t)(ctx->launches - launches_before);
        ctx->launches = launches_before;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        if (rc != CB200_OK || e != cudaSuccess || !graph) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            ctx->use_graph = false;
            return rc;
        }
        slot = &ctx->graphs[ctx->graph_next];
        ctx->graph_next = (ctx->graph_next + 1) % cb200_ctx::kGraphSlots;
        if (slot->exec) cudaGraphExecDestroy(slot->exec);
        slot->exec = nullptr;
        const cudaError_t ei = cudaGraphInstantiate(&slot->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) {
            cudaGetLastError();
            slot->exec = nullptr;
            ctx->use_graph = false;
            return CB200_OK;
        }
        slot->sig = sig;
        ctx->graph_kernels[slot - ctx->graphs] = n_kernels;
        ctx->graph_captures += 1;
        if (getenv("CB200_VERBOSE")) fprintf(stderr, "[cb200] captured step graph #%llu (%u kernels), signature %016llx\n", (unsigned long long)ctx->graph_captures, n_kernels, (unsigned long long)sig);
    }
    CU(cudaGraphLaunch(slot->exec, ctx->stream));
    ctx->launches += ctx->graph_kernels[slot - ctx->graphs];
    CU(cudaEventRecord(ctx->ev_t[ST_COUNT], ctx->stream));
    ctx->graph_launches += 1;
    *replayed = true;
    return CB200_OK;
}

static int flush_deferred_fetch(cb200_ctx* ctx);
int cb200_bulk_update(cb200_ctx* ctx, double time, const cb200_vel_view* vel, double end_time, cb200_step_result* out) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->have_state) return fail(ctx, CB200_E_STATE, "bulk_update before upload_state");
    if (ctx->sharded) return fail(ctx, CB200_E_STATE, "context is sharded: use cb200_shard_begin / cb200_shard_finish");
    if (!(time < 1e50)) return fail(ctx, CB200_E_ARG, "requires time < 1e50");  // collider.rs:100
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev_t[ST_PRE], ctx->stream));
    const double* nv[5];
    int rc = stage_vel(ctx, vel, nv);
    if (rc != CB200_OK) return rc;
    if ((rc = prepare_buffers(ctx)) != CB200_OK) return rc;
    if (ctx->ov_track && !ctx->ov_ready && (rc = build_overlap_tables(ctx, ctx->n_ov)) != CB200_OK) return rc;
    if (ctx->ov_track && vel && vel->end_time) return fail(ctx, CB200_E_ARG, "the window drain needs one end time for the whole step (vel->end_time must be NULL)");
    ctx->tile_steps += 1;
    if ((rc = set_step_dyn(ctx, time, end_time)) != CB200_OK) return rc;
    // A step is re-run (without the rebase) when a buffer had to grow — or, on the tile path, when a tile exceeded its
    // shared-memory capacity (then on the global path).
    bool again = true;
    for (int attempt = 0; attempt < 5 && again; ++attempt) {
        ctx->step_is_tile = ctx->tile_enabled && ctx->tile_ok && !ctx->tile_suspended && !ctx->add_mode && ctx->n >= 2;
        bool replayed = false;
        if (attempt == 0 && ctx->use_graph && !ctx->stage_timing) {
            if ((rc = launch_step_graph(ctx, time, nv, end_time, &replayed)) != CB200_OK) return rc;
        }
        if (!replayed) {
            if (ctx->step_is_tile) {
                if ((rc = enqueue_tile_step(ctx, attempt == 0, nv)) != CB200_OK) return rc;
            } else {
                if ((rc = enqueue_front(ctx, time, attempt == 0, nv, end_time)) != CB200_OK) return rc;
                if ((rc = enqueue_back(ctx, time, false)) != CB200_OK) return rc;
            }
        }
        if ((rc = flush_deferred_fetch(ctx)) != CB200_OK) return rc;  // (the previous step's event download: behind this step's launch)
        if ((rc = complete_step(ctx, time, &again)) != CB200_OK) return rc;
        if (ctx->last.ctr.err_flags) break;
    }
    if (again && !ctx->last.ctr.err_flags) return fail(ctx, CB200_E_STATE, "step buffers failed to converge");
    accumulate_timing(ctx, ST_PRE, true);
    ctx->timed_steps += 1;
    ctx->timing_valid = true;
    return fill_result(ctx, out);
}

// ---- sharded step -------------------------------------------------------------------------
int cb200_shard_configure(cb200_ctx* ctx, int rank, int world, const int32_t* col_bounds, uint64_t id_space) {
    if (!ctx || !col_bounds) return CB200_E_ARG;
    if (id_space == 0 || id_space >= 0x7fffffffull) return fail(ctx, CB200_E_ARG, "id_space must be in 1 .. 2^31 - 2 (ids of all ranks are < id_space)");
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return fail(ctx, CB200_E_ARG, "bad rank / world");
    for (int d = 0; d < world; ++d)
        if (!(col_bounds[d] < col_bounds[d + 1])) return fail(ctx, CB200_E_ARG, "strip bounds must be strictly ascending");
    ShardGeom g{};
    g.world = world;
    g.rank = rank;
    for (int d = 0; d <= world; ++d) g.bounds[d] = col_bounds[d];
    g.bounds[0] = INT32_MIN;
    g.bounds[world] = INT32_MAX;
    g.col_lo = g.bounds[rank];
    g.col_hi = g.bounds[rank + 1];
    ctx->shard = g;
    ctx->id_space = id_space;
    ctx->sharded = true;
    ctx->have_state = false;  // the table must be (re-)uploaded: sharded tables carry label = id
    return CB200_OK;
}

int cb200_shard_set_slab(cb200_ctx* ctx, uint64_t records) {
    if (!ctx) return CB200_E_ARG;
    if (records < 1 || records >= 0x7ffffff0ull) return fail(ctx, CB200_E_ARG, "bad slab size");
    ctx->slab = (uint32_t)records + 1;
    ctx->have_state = false;  // buffers are sized at upload_state
    return CB200_OK;
}

int cb200_shard_set_row_capacity(cb200_ctx* ctx, uint64_t rows) {
    if (!ctx) return CB200_E_ARG;
    if (rows >= 0x7fffffffull) return fail(ctx, CB200_E_ARG, "bad row capacity");
    ctx->own_cap_request = (uint32_t)rows;  // 0 = default (residents + 25 %)
    ctx->have_state = false;                // buffers are sized at upload_state
    return CB200_OK;
}

int cb200_shard_rows(const cb200_ctx* ctx, uint64_t* rows, uint64_t* capacity, uint64_t* migrated_out, uint64_t* migrated_in) {
    if (!ctx) return CB200_E_ARG;
    if (rows) *rows = ctx->n;
    if (capacity) *capacity = ctx->sharded ? ctx->own_cap : ctx->n;
    if (migrated_out) *migrated_out = ctx->rows_migrated_out;
    if (migrated_in) *migrated_in = ctx->rows_migrated_in;
    return CB200_OK;
}

static MigCols mig_cols_of(cb200_ctx* ctx) {
    MigCols c{};
    for (int k = 0; k < 11; ++k) c.col[k] = ctx->col[k].p;
    c.kind = ctx->kind.p; c.reit = ctx->reit.p; c.group = ctx->group.p; c.mask = ctx->mask.p; c.label = ctx->label.p; c.ord = ctx->ord.p;
    return c;
}
static int mig_check(cb200_ctx* ctx, const char* what) {
    if (!ctx->sharded) return fail(ctx, CB200_E_STATE, std::string(what) + ": cb200_shard_configure first");
    if (!ctx->have_state) return fail(ctx, CB200_E_STATE, std::string(what) + " before upload_state");
    if (ctx->shard_phase != 0) return fail(ctx, CB200_E_STATE, std::string(what) + " inside a step (cb200_shard_result first)");
    if (ctx->fetch_pending) return fail(ctx, CB200_E_STATE, std::string(what) + ": cb200_fetch_events_end the pending download first");
    return CB200_OK;
}
static void mig_invalidate(cb200_ctx* ctx) {
    ctx->have_step = false;  // the event list of the last step named rows that may be gone
    ctx->ev_sorted = false;
    ctx->tile_ok = false;
    set_l2_window(ctx);
}

// Owner migration, sending side (migrate.cuh): the rows whose centre column lies more than margin_cols columns outside this
// rank's strip are copied out — every column, plus the rank whose strip holds them — and removed from the table.
int cb200_shard_take_movers(cb200_ctx* ctx, int32_t margin_cols, const cb200_rows_out* out, uint64_t* n_out) {
    if (!ctx) return CB200_E_ARG;
    if (n_out) *n_out = 0;
    int rc = mig_check(ctx, "shard_take_movers");
    if (rc != CB200_OK) return rc;
    if (margin_cols < 0) return fail(ctx, CB200_E_ARG, "margin_cols must be >= 0");
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->n;
    if (!n) return CB200_OK;
    CU(ctx->sort_key.reserve(n)); CU(ctx->src_row.reserve(n)); CU(ctx->mig_u32.reserve(4096));
    uint32_t* rows = ctx->sort_key.p;
    int32_t* dest = reinterpret_cast<int32_t*>(ctx->src_row.p);
    uint32_t* count = ctx->mig_u32.p;
    const uint32_t blocks = (n + kThreads - 1) / kThreads;
    CU(cudaMemsetAsync(count, 0, 4, ctx->stream));
    k_mig_mark<<<blocks, kThreads, 0, ctx->stream>>>(n, ctx->col[0].p, ctx->cell_width, ctx->shard, margin_cols, rows, dest, count, n);
    ctx->launches += 1;
    uint32_t m = 0;
    CU(cudaMemcpyAsync(&m, count, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (n_out) *n_out = m;
    if (!m) return CB200_OK;
    if (!out || out->cap < m) return fail(ctx, CB200_E_ARG, "shard_take_movers: output capacity too small (nothing was removed; *n_out = rows that would move)");
    const void* need[] = {out->id, out->pos_x, out->pos_y, out->dim_x, out->dim_y, out->vel_x, out->vel_y, out->rsz_x, out->rsz_y, out->start_time,
                          out->end_time, out->pub_end_time, out->kind, out->group, out->interact_mask, out->dest};
    for (const void* q : need)
        if (!q) return fail(ctx, CB200_E_ARG, "null column in rows_out");
    // pack + download
    const size_t bytes = mig_stage_bytes(m);
    CU(ctx->mig_stage.reserve(bytes));
    const MigCols mc = mig_cols_of(ctx);
    const uint32_t mblocks = (m + kThreads - 1) / kThreads;
    k_mig_pack<<<mblocks, kThreads, 0, ctx->stream>>>(m, rows, mc, ctx->mig_stage.p, ctx->vmap.p, (uint32_t)ctx->id_space);
    ctx->launches += 1;
    std::vector<unsigned char> hs(bytes);
    std::vector<uint32_t> hrows(m);
    CU(cudaMemcpyAsync(hs.data(), ctx->mig_stage.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(hrows.data(), rows, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out->dest, dest, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    double* oc[11] = {out->pos_x, out->pos_y, out->dim_x, out->dim_y, out->vel_x, out->vel_y, out->rsz_x, out->rsz_y, out->start_time, out->end_time, out->pub_end_time};
    const double* hd = reinterpret_cast<const double*>(hs.data());
    for (int k = 0; k < 11; ++k) memcpy(oc[k], hd + (size_t)k * m, (size_t)m * 8);
    const uint32_t* hu = reinterpret_cast<const uint32_t*>(hd + (size_t)11 * m);
    for (uint32_t j = 0; j < m; ++j) out->id[j] = hu[j];
    std::vector<uint32_t> gone(hu + m, hu + 2 * (size_t)m);  // the movers' positions in the caller's array order
    memcpy(out->group, hu + 2 * (size_t)m, (size_t)m * 4);
    memcpy(out->interact_mask, hu + 3 * (size_t)m, (size_t)m * 4);
    memcpy(out->kind, reinterpret_cast<const unsigned char*>(hu + 4 * (size_t)m), m);
    // holes among the first n - m rows <- the surviving rows of the tail
    const uint32_t n_left = n - m;
    std::sort(hrows.begin(), hrows.end());
    std::sort(gone.begin(), gone.end());
    std::vector<uint32_t> up;  // [moves: (dst, src) pairs][gone ords]
    {
        size_t t = std::lower_bound(hrows.begin(), hrows.end(), n_left) - hrows.begin();  // movers inside the tail start here
        uint32_t src = n_left;
        for (size_t h = 0; h < hrows.size() && hrows[h] < n_left; ++h) {
            while (t < hrows.size() && hrows[t] == src) { ++t; ++src; }  // (skip tail rows that move away themselves)
            up.push_back(hrows[h]);
            up.push_back(src++);
        }
    }
    const uint32_t n_moves = (uint32_t)(up.size() / 2);
    up.insert(up.end(), gone.begin(), gone.end());
    CU(ctx->mig_u32.reserve(up.size() + 16));
    CU(cudaMemcpyAsync(ctx->mig_u32.p, up.data(), up.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (n_moves) k_mig_fill<<<(n_moves + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(n_moves, ctx->mig_u32.p, mc);
    if (n_left) k_mig_ord_out<<<(n_left + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(n_left, ctx->ord.p, ctx->mig_u32.p + 2 * (size_t)n_moves, m);
    ctx->launches += 2;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->n = n_left;
    ctx->rows_migrated_out += m;
    mig_invalidate(ctx);
    return CB200_OK;
}

// Owner migration, receiving side: the rows are appended to the table (ids strictly ascending, none of them resident here).
int cb200_shard_add_rows(cb200_ctx* ctx, const cb200_state_view* v) {
    if (!ctx || !v) return CB200_E_ARG;
    int rc = mig_check(ctx, "shard_add_rows");
    if (rc != CB200_OK) return rc;
    const uint64_t m64 = v->n;
    if (!m64) return CB200_OK;
    const void* need[] = {v->id, v->pos_x, v->pos_y, v->dim_x, v->dim_y, v->vel_x, v->vel_y, v->rsz_x, v->rsz_y,
                          v->start_time, v->end_time, v->pub_end_time, v->kind, v->group, v->interact_mask};
    for (const void* q : need)
        if (!q) return fail(ctx, CB200_E_ARG, "null column in state view");
    const uint32_t n = ctx->n;
    if ((uint64_t)n + m64 > ctx->own_cap) return fail(ctx, CB200_E_STATE, "shard_add_rows: the strip's row capacity is exhausted (cb200_shard_set_row_capacity before upload_state)");
    const uint32_t m = (uint32_t)m64;
    std::vector<uint32_t> up(m);
    for (uint32_t j = 0; j < m; ++j) {
        if (j && !(v->id[j - 1] < v->id[j])) return fail(ctx, CB200_E_ARG, "ids must be strictly ascending");
        if (v->id[j] >= ctx->id_space) return fail(ctx, CB200_E_ARG, "sharded mode requires ids < id_space");
        if (v->group[j] != CB200_GROUP_NONE && v->group[j] > 31) return fail(ctx, CB200_E_ARG, "group must be 0..31 or CB200_GROUP_NONE");
        up[j] = (uint32_t)v->id[j];
    }
    CU(cudaSetDevice(ctx->device));
    const double* src[11] = {v->pos_x, v->pos_y, v->dim_x, v->dim_y, v->vel_x, v->vel_y, v->rsz_x, v->rsz_y, v->start_time, v->end_time, v->pub_end_time};
    for (int k = 0; k < 11; ++k) CU(cudaMemcpyAsync(ctx->col[k].p + n, src[k], (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->kind.p + n, v->kind, m, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->group.p + n, v->group, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->mask.p + n, v->interact_mask, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->label.p + n, up.data(), (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->reit.p + n, 0, m, ctx->stream));
    // the caller's array order is the ascending-id order of the residents: shift the residents' positions, place the arrivals
    CU(ctx->mig_u32.reserve(2 * (size_t)m + 16));
    uint32_t* d_ids = ctx->mig_u32.p;
    uint32_t* d_below = ctx->mig_u32.p + m;  // [m + 1]
    CU(cudaMemcpyAsync(d_ids, up.data(), (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(d_below, 0, ((size_t)m + 1) * 4, ctx->stream));
    if (n) k_mig_ord_in<<<(n + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(n, ctx->ord.p, ctx->label.p, d_ids, m, d_below);
    ctx->launches += 1;
    std::vector<uint32_t> below((size_t)m + 1);
    CU(cudaMemcpyAsync(below.data(), d_below, ((size_t)m + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    uint32_t run = 0;
    for (uint32_t j = 0; j < m; ++j) {
        run += below[j];
        up[j] = run + j;
    }
    CU(cudaMemcpy(ctx->ord.p + n, up.data(), (size_t)m * 4, cudaMemcpyHostToDevice));
    ctx->n = n + m;
    ctx->rows_migrated_in += m;
    if ((uint64_t)m * 16 > ctx->n) ctx->need_resort = true;  // (many new rows at the end of the table: put them into slot order)
    mig_invalidate(ctx);
    return CB200_OK;
}

int cb200_set_stream(cb200_ctx* ctx, void* cuda_stream) {
    if (!ctx) return CB200_E_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return CB200_OK;
}

// Asynchronous: enqueues stage A + halo pack on the context's stream and returns the buffers of the exchange.
int cb200_shard_begin(cb200_ctx* ctx, double time, const cb200_vel_view* vel, double end_time, void** send_base, void** recv_base,
                      uint64_t* slab_bytes) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->sharded) return fail(ctx, CB200_E_STATE, "cb200_shard_configure first");
    if (!ctx->have_state) return fail(ctx, CB200_E_STATE, "shard_begin before upload_state");
    if (!(time < 1e50)) return fail(ctx, CB200_E_ARG, "requires time < 1e50");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev_t[ST_PRE], ctx->stream));
    const double* nv[5];
    int rc = stage_vel(ctx, vel, nv);
    if (rc != CB200_OK) return rc;
    if ((rc = prepare_buffers(ctx)) != CB200_OK) return rc;
    if ((rc = set_step_dyn(ctx, time, end_time)) != CB200_OK) return rc;
    if ((rc = enqueue_front(ctx, time, true, nv, end_time)) != CB200_OK) return rc;
    ctx->shard_phase = 1;
    ctx->shard_T = time;
    if (send_base) *send_base = ctx->send.p;
    if (recv_base) *recv_base = ctx->rec.p + ctx->own_cap;
    if (slab_bytes) *slab_bytes = (uint64_t)ctx->slab * sizeof(HbRec);
    return CB200_OK;
}

// Asynchronous: enqueues the back half (the exchange must already be ordered before it on the same stream).
// *local_key receives the device address of this rank's (time key, tie) minimum — 16 bytes, the all-gather input.
int cb200_shard_finish(cb200_ctx* ctx, void** local_key) {
    if (!ctx) return CB200_E_ARG;
    if (ctx->shard_phase != 1) return fail(ctx, CB200_E_STATE, "shard_finish without shard_begin");
    CU(cudaSetDevice(ctx->device));
    int rc = enqueue_back(ctx, ctx->shard_T, false);
    if (rc != CB200_OK) return rc;
    ctx->shard_phase = 2;
    if (local_key) *local_key = ctx->d_key;
    return flush_deferred_fetch(ctx);
}

// Synchronous: waits for the step; re-runs the back half if a buffer had to grow; fills the (local) result.
int cb200_shard_result(cb200_ctx* ctx, cb200_step_result* out) {
    if (!ctx) return CB200_E_ARG;
    if (ctx->shard_phase != 2) return fail(ctx, CB200_E_STATE, "shard_result without shard_finish");
    CU(cudaSetDevice(ctx->device));
    ctx->shard_phase = 0;
    int rc;
    bool again = false;
    if ((rc = complete_step(ctx, ctx->shard_T, &again)) != CB200_OK) {
        cb200_shard_abort(ctx);
        return rc;
    }
    if (ctx->last.ctr.overflow_send) {
        cb200_shard_abort(ctx);
        return fail(ctx, CB200_E_STATE, "halo slab overflow: more records leave this rank's strip than the slab holds (cb200_shard_set_slab)");
    }
    accumulate_timing(ctx, ST_PRE, true);
    for (int attempt = 0; attempt < 3 && again && !ctx->last.ctr.err_flags; ++attempt) {
        // re-run of the back half only: slot table and counters are rebuilt, the visitor table is intact
        const uint32_t n_slots = ctx->geom.n_slots();
        const Counters keep = ctx->last.ctr;
        Counters z{};
        z.err_slot = keep.err_slot;
        z.err_flags = keep.err_flags;
        z.n_solitaire = keep.n_solitaire;
        z.n_local = keep.n_local;
        CU(cudaMemsetAsync(ctx->slots.p, 0, (size_t)n_slots * 4, ctx->stream));
        CU(cudaMemsetAsync(ctx->work_count.p, 0, (kWorkLists + 2) * 4, ctx->stream));
        CU(cudaMemcpy(ctx->d_ctr, &z, sizeof(Counters), cudaMemcpyHostToDevice));
        if ((rc = enqueue_back(ctx, ctx->shard_T, true)) != CB200_OK) return rc;
        if ((rc = complete_step(ctx, ctx->shard_T, &again)) != CB200_OK) return rc;
    }
    if (again && !ctx->last.ctr.err_flags) return fail(ctx, CB200_E_STATE, "step buffers failed to converge");
    ctx->timed_steps += 1;
    ctx->timing_valid = true;
    rc = fill_result(ctx, out);
    if (rc != CB200_OK) cb200_shard_abort(ctx);  // this rank stops stepping: do not leave the peers waiting for it
    return rc;
}

// Tell every peer of a direct-exchange rank that this rank gives up (their pending waits end with CB200_F_PEER_LOST).
// Called by the library when a sharded step fails; a host that catches an error of its own between steps calls it too.
int cb200_shard_abort(cb200_ctx* ctx) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->p2p || ctx->shard.world < 2) return CB200_OK;
    cudaSetDevice(ctx->device);
    k_raise_abort<<<1, kMaxWorld, 0, ctx->stream>>>(ctx->peers, ctx->shard.world, ctx->shard.rank);
    ctx->launches += 1;
    cudaStreamSynchronize(ctx->stream);
    cudaGetLastError();
    return CB200_OK;
}

// ---- direct exchange over peer memory --------------------------------------------------------------------
int cb200_shard_p2p_export(cb200_ctx* ctx, cb200_p2p_info* out) {
    if (!ctx || !out) return CB200_E_ARG;
    if (!ctx->sharded || !ctx->have_state) return fail(ctx, CB200_E_STATE, "p2p_export needs a configured, uploaded sharded context");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->d_mail) {
        CU(cudaMalloc(&ctx->d_mail, sizeof(Mailbox)));
        CU(cudaHostAlloc(&ctx->h_gkey, sizeof(GlobalKey), cudaHostAllocMapped));
        CU(cudaHostGetDevicePointer((void**)&ctx->d_gkey, ctx->h_gkey, 0));
    }
    CU(cudaMemset(ctx->d_mail, 0, sizeof(Mailbox)));
    memset(out, 0, sizeof(*out));
    static_assert(sizeof(cudaIpcMemHandle_t) == CB200_IPC_HANDLE_BYTES, "ipc handle size");
    cudaIpcMemHandle_t h;
    // (a failing cudaIpcGetMemHandle is not fatal: peers in the same process use the raw pointers)
    if (cudaIpcGetMemHandle(&h, ctx->rec.p) == cudaSuccess) memcpy(out->rec_handle, &h, sizeof(h));
    else cudaGetLastError();
    if (cudaIpcGetMemHandle(&h, ctx->d_mail) == cudaSuccess) memcpy(out->mail_handle, &h, sizeof(h));
    else cudaGetLastError();
    out->rec_ptr = (uint64_t)(uintptr_t)ctx->rec.p;
    out->mail_ptr = (uint64_t)(uintptr_t)ctx->d_mail;
    out->recv_offset_bytes = (uint64_t)ctx->own_cap * sizeof(HbRec);
    out->pid = (uint64_t)getpid();
    out->device = ctx->device;
    out->slab_records = ctx->slab;
    return CB200_OK;
}

int cb200_shard_p2p_import(cb200_ctx* ctx, const cb200_p2p_info* infos /* world entries, indexed by rank */) {
    if (!ctx || !infos) return CB200_E_ARG;
    if (!ctx->sharded || !ctx->d_mail) return fail(ctx, CB200_E_STATE, "p2p_import after p2p_export");
    CU(cudaSetDevice(ctx->device));
    for (int i = 0; i < ctx->n_ipc_opened; ++i) cudaIpcCloseMemHandle(ctx->ipc_opened[i]);
    ctx->n_ipc_opened = 0;
    const int world = ctx->shard.world, rank = ctx->shard.rank;
    ctx->p2p_same_process = false;
    PeerTable pt{};
    for (int d = 0; d < world; ++d) {
        if (infos[d].slab_records != ctx->slab) return fail(ctx, CB200_E_ARG, "peers must use the same halo slab size (cb200_shard_set_slab)");
        if (d == rank) {
            pt.recv[d] = ctx->rec.p + ctx->own_cap;
            pt.mail[d] = ctx->d_mail;
            continue;
        }
        char* rec_base = nullptr;
        Mailbox* mail = nullptr;
        if (infos[d].pid == (uint64_t)getpid()) {
            ctx->p2p_same_process = true;
            // same process (several contexts of one host, e.g. the tests): the peer's pointers are valid here
            rec_base = (char*)(uintptr_t)infos[d].rec_ptr;
            mail = (Mailbox*)(uintptr_t)infos[d].mail_ptr;
            if (infos[d].device != ctx->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(infos[d].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, CB200_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
        } else {
            cudaIpcMemHandle_t h;
            void* p = nullptr;
            memcpy(&h, infos[d].rec_handle, sizeof(h));
            CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            ctx->ipc_opened[ctx->n_ipc_opened++] = p;
            rec_base = (char*)p;
            memcpy(&h, infos[d].mail_handle, sizeof(h));
            CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            ctx->ipc_opened[ctx->n_ipc_opened++] = p;
            mail = (Mailbox*)p;
        }
        pt.recv[d] = reinterpret_cast<HbRec*>(rec_base + infos[d].recv_offset_bytes);
        pt.mail[d] = mail;
    }
    ctx->peers = pt;
    ctx->p2p = true;
    ctx->p2p_step = 0;
    // CUDA loads a kernel's code at its first launch, and that load waits for the device to go idle: it must not happen
    // behind a kernel that is waiting for a peer.  Touch every kernel of the step now.
    {
        cudaFuncAttributes fa;
        const void* fns[] = {(const void*)k_step_zero, (const void*)k_stage_a<true>, (const void*)k_stage_a_tma<true>, (const void*)k_slab_header, (const void*)k_push_halo,
                             (const void*)k_wait_flags, (const void*)k_index_visitors<true>, (const void*)k_index_visitors<false>,
                             (const void*)k_scan_slots, (const void*)k_scatter<true>, (const void*)k_enum_pairs, (const void*)k_solve<true, true>, (const void*)k_solve<true, false>, (const void*)k_solve<false, true>, (const void*)k_solve<false, false>, (const void*)k_solve2, (const void*)k_solve_dense, (const void*)k_solve_dense_lockstep,
                             (const void*)k_push_key, (const void*)k_reduce_keys, (const void*)k_raise_abort, (const void*)k_resort_count, (const void*)k_resort_rank,
                             (const void*)k_resort_gather};
        for (const void* f : fns) CU(cudaFuncGetAttributes(&fa, f));
    }
    return reserve_step_buffers(ctx);
}

// One whole sharded step, asynchronous on the context's stream: stage A + halo pack, push the slab to every peer, wait for
// theirs, bin / enumerate / solve, exchange the next-event keys.  cb200_shard_result waits for it.
int cb200_shard_step(cb200_ctx* ctx, double time, const cb200_vel_view* vel, double end_time) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->p2p) return fail(ctx, CB200_E_STATE, "cb200_shard_p2p_import first");
    if (!ctx->have_state) return fail(ctx, CB200_E_STATE, "shard_step before upload_state");
    if (!(time < 1e50)) return fail(ctx, CB200_E_ARG, "requires time < 1e50");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev_t[ST_PRE], ctx->stream));
    const double* nv[5];
    int rc = stage_vel(ctx, vel, nv);
    if (rc != CB200_OK) return rc;
    if ((rc = prepare_buffers(ctx)) != CB200_OK) return rc;
    ctx->p2p_step += 1;
    ctx->p2p_parity = (int)(ctx->p2p_step & 1);
    if ((rc = set_step_dyn(ctx, time, end_time)) != CB200_OK) return rc;
    ctx->shard_T = time;
    bool replayed = false;
    // (graph instantiation may synchronise the device: not while a peer context of this very process is waiting in a kernel)
    if (ctx->use_graph && !ctx->stage_timing && !ctx->p2p_same_process) {
        if ((rc = launch_step_graph(ctx, time, nv, end_time, &replayed)) != CB200_OK) return rc;
    }
    if (!replayed && (rc = enqueue_direct_step(ctx, time, nv, end_time)) != CB200_OK) return rc;
    CU(cudaEventRecord(ctx->ev_t[ST_COUNT], ctx->stream));
    ctx->shard_phase = 2;
    return flush_deferred_fetch(ctx);  // (the previous step's event download: behind this step's launch)
}

// After cb200_shard_result of a direct-exchange step: the global next event (minimum over all ranks).  *complete == 0:
// some rank had to re-run the step with larger buffers after the keys were exchanged — all-gather the local keys again
// (cb200_shard_finish's local_key) to get the final minimum.
int cb200_shard_global_next(cb200_ctx* ctx, uint64_t* time_key_out, uint64_t* tie_out, int* complete) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->p2p || !ctx->h_gkey || ctx->h_gkey->tag != ctx->p2p_step) return fail(ctx, CB200_E_STATE, "no direct-exchange step has completed");
    if (ctx->h_gkey->lost) return fail(ctx, CB200_E_CONTRACT, flags_message(kFPeerLost));
    if (time_key_out) *time_key_out = ctx->h_gkey->t;
    if (tie_out) *tie_out = ctx->h_gkey->tie;
    if (complete) *complete = (int)ctx->h_gkey->complete;
    return CB200_OK;
}

void* cb200_shard_local_key(cb200_ctx* ctx) { return ctx ? ctx->d_key : nullptr; }

// ---- event queue materialisation -------------------------------------------------------------------------
// Sorts keys[0, n) in place (scratch: the same range of `tmp`) by the bucket sort of sort.cuh.  field: 1 = cut the buckets on
// the time key, 0 = on the tie word (all keys of the range share one time).  *ok = false: give up (the caller falls back
// to the radix sort).  One small D2H + sync per level.
static int sort_range(cb200_ctx* ctx, Key128* keys, Key128* tmp, uint32_t n, int field, int depth, bool* ok) {
    *ok = false;
    if (n < 2) {
        *ok = true;
        return CB200_OK;
    }
    if (depth > 3) return CB200_OK;
    int log2_buckets = 6;
    while ((1u << log2_buckets) * 12u < n && log2_buckets < 22) ++log2_buckets;
    const uint32_t n_buckets = 1u << log2_buckets;
    const size_t words = 2 * (size_t)n_buckets + 2 + 2 * kBsMaxBig + 8;
    DevBuf<uint32_t>& tab = ctx->bs_tab[depth];  // (deeper levels run while the parent's table is still needed: one per depth)
    CU(tab.reserve(words));
    uint32_t* count = tab.p;                      // then: offsets [n_buckets + 1]
    uint32_t* cursor = tab.p + n_buckets + 1;     // [n_buckets]
    uint32_t* big = cursor + n_buckets;           // [1 + 2 * kBsMaxBig]
    BsRange* range = reinterpret_cast<BsRange*>(tab.p + ((2 * (size_t)n_buckets + 2 + 2 * kBsMaxBig + 1) & ~(size_t)1));
    const BsRange init{~0ull, 0ull};
    CU(cudaMemsetAsync(count, 0, ((size_t)n_buckets + 1) * 4, ctx->stream));
    CU(cudaMemcpyAsync(range, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t blocks = (n + kThreads - 1) / kThreads;
    k_bs_minmax<<<std::min<uint32_t>(blocks, 1024), kThreads, 0, ctx->stream>>>(n, keys, field, range);
    ctx->launches += 1;
    if (depth > 0 && field == 1) {
        // an oversized bucket of the time key: if all its times are equal, the tie word decides
        BsRange r;
        CU(cudaMemcpyAsync(&r, range, sizeof(r), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (r.lo == r.hi) return sort_range(ctx, keys, tmp, n, 0, depth, ok);
    }
    k_bs_count<<<blocks, kThreads, 0, ctx->stream>>>(n, keys, field, range, log2_buckets, count);
    k_bs_scan<<<1, 1024, 0, ctx->stream>>>(count, n_buckets, count, cursor, big);
    k_bs_scatter<<<blocks, kThreads, 0, ctx->stream>>>(n, keys, field, range, log2_buckets, cursor, tmp);
    k_bs_local<<<std::min<uint32_t>((n_buckets + 7) / 8, (uint32_t)ctx->sm_count * 8u), 256, 0, ctx->stream>>>(tmp, count, n_buckets, keys);
    ctx->launches += 4;
    uint32_t h_big[1 + 2 * kBsMaxBig];
    CU(cudaMemcpyAsync(h_big, big, sizeof(h_big), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (getenv("CB200_VERBOSE")) fprintf(stderr, "[cb200] event sort level %d field %d: %u keys, %u buckets, %u oversized\n", depth, field, n, n_buckets, h_big[0]);
    if (h_big[0] > (uint32_t)kBsMaxBig) {
        // too fragmented: fall back.  The oversized buckets exist only in tmp: hand the complete multiset back first.
        CU(cudaMemcpyAsync(keys, tmp, (size_t)n * sizeof(Key128), cudaMemcpyDeviceToDevice, ctx->stream));
        return CB200_OK;
    }
    for (uint32_t k = 0; k < h_big[0]; ++k) {
        const uint32_t off = h_big[1 + 2 * k], m = h_big[2 + 2 * k];
        if (m == n) {
            // no progress: every key has the same value in this field (the keys are in tmp, unsorted)
            CU(cudaMemcpyAsync(keys, tmp, (size_t)n * sizeof(Key128), cudaMemcpyDeviceToDevice, ctx->stream));
            if (field == 1) return sort_range(ctx, keys, tmp, n, 0, depth + 1, ok);
            return CB200_OK;
        }
        // the bucket's keys sit in tmp (k_bs_local skipped it): sort them there, then bring them over
        bool sub_ok = false;
        int rc = sort_range(ctx, tmp + off, keys + off, m, field, depth + 1, &sub_ok);
        if (rc != CB200_OK) return rc;
        CU(cudaMemcpyAsync(keys + off, tmp + off, (size_t)m * sizeof(Key128), cudaMemcpyDeviceToDevice, ctx->stream));
        if (!sub_ok) {
            // give up: the remaining oversized buckets are still only in tmp
            for (uint32_t k2 = k + 1; k2 < h_big[0]; ++k2)
                CU(cudaMemcpyAsync(keys + h_big[1 + 2 * k2], tmp + h_big[1 + 2 * k2], (size_t)h_big[2 + 2 * k2] * sizeof(Key128), cudaMemcpyDeviceToDevice, ctx->stream));
            return CB200_OK;
        }
    }
    *ok = true;
    return CB200_OK;
}

// Order keys of every queued event of the last step, in keys_a (queued on the stream); range (optional, initialised by the
// caller): min / max of their time keys.
static int build_sort_keys(cb200_ctx* ctx, uint32_t n_pairs, uint64_t n_solitaire, BsRange* range, BsRange* lab_range = nullptr) {
    CU(cudaMemsetAsync(&ctx->d_ctr->n_sort_keys, 0, 8, ctx->stream));
    const uint32_t max_blocks = (uint32_t)ctx->sm_count * 8u;
    if (ctx->n && n_solitaire)
        k_keys_solitaire<<<std::min<uint32_t>((ctx->n + kThreads - 1) / kThreads, max_blocks), kThreads, 0, ctx->stream>>>(ctx->n, ctx->reit.p, ctx->col[9].p, ctx->label.p,
                                                                                                                     ctx->keys_a.p, ctx->d_ctr, range, lab_range);
    if (n_pairs)
        k_keys_pairs<<<std::min<uint32_t>((n_pairs + kThreads - 1) / kThreads, max_blocks), kThreads, 0, ctx->stream>>>(n_pairs, ctx->events.p, ctx->keys_a.p, n_solitaire, range, lab_range);
    ctx->launches += (ctx->n && n_solitaire ? 1 : 0) + (n_pairs ? 1 : 0);
    return CB200_OK;
}

// One-shot bucket sort (sort.cuh): everything is queued without a host round trip — keys, two-region bucket cut, rank sort,
// decode into ev_out.  Whether it succeeded (no bucket too large for a warp) is in ctx->h_sort_big[0] once the stream has
// been synchronised; the caller checks it and falls back to sort_events_slow.
// `st`: the stream the sort itself runs on.  The order keys are always built on the context's stream (they read the step's
// event buffer and state columns, which the next step overwrites); with st != ctx->stream the bucket sort then runs beside
// whatever the context's stream does next (cb200_fetch_events_begin: beside the next bulk step).  decode32: also decode the
// sorted keys into the 32-byte cb200_event records (ev_out).
static int enqueue_sort_oneshot(cb200_ctx* ctx, uint32_t nt, uint32_t n_pairs, uint64_t n_solitaire, cudaStream_t st, bool decode32) {
    int log2_buckets = 6;
    while ((1u << log2_buckets) * 8u < nt && log2_buckets < 24) ++log2_buckets;  // ~8 keys per bucket (a warp sorts up to 256)
    const uint32_t nb = 1u << log2_buckets, n_buckets = 2 * nb;  // region A (ties at the first time) + region B (by time)
    const uint32_t scan_blocks = (n_buckets + kBs2Chunk - 1) / kBs2Chunk;
    DevBuf<uint32_t>& tab = ctx->bs_tab[0];
    CU(tab.reserve(2 * (size_t)n_buckets + 2 + 12 + scan_blocks));
    uint32_t* count = tab.p;                   // [n_buckets + 1] then: offsets
    uint32_t* cursor = tab.p + n_buckets + 4;  // [n_buckets] (16-byte aligned like count)
    BsRange* range = reinterpret_cast<BsRange*>(tab.p + 2 * (size_t)n_buckets + 4);  // [2]: time keys, first labels
    uint32_t* block_total = tab.p + 2 * (size_t)n_buckets + 12;
    static const BsRange init[2] = {{~0ull, 0ull}, {~0ull, 0ull}};
    CU(cudaMemsetAsync(count, 0, ((size_t)n_buckets + 1) * 4, ctx->stream));
    CU(cudaMemcpyAsync(range, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    int rc = build_sort_keys(ctx, n_pairs, n_solitaire, range, range + 1);
    if (rc != CB200_OK) return rc;
    if (st != ctx->stream) {
        CU(cudaEventRecord(ctx->ev_keys_done, ctx->stream));
        CU(cudaStreamWaitEvent(st, ctx->ev_keys_done, 0));
    }
    const uint32_t blocks = (nt + kThreads - 1) / kThreads;
    Key128 *keys = ctx->keys_a.p, *tmp = ctx->keys_b.p;
    k_bs2_count<<<blocks, kThreads, 0, st>>>(nt, keys, range, log2_buckets, count);
    k_bs2_scan_local<<<scan_blocks, 256, 0, st>>>(count, n_buckets, cursor, block_total);
    k_bs2_scan_add<<<scan_blocks, 256, 0, st>>>(count, n_buckets, cursor, block_total, ctx->d_sort_big);
    k_bs2_scatter<<<blocks, kThreads, 0, st>>>(nt, keys, range, log2_buckets, cursor, tmp);
    k_bs_local<<<std::min<uint32_t>((n_buckets + 7) / 8, (uint32_t)ctx->sm_count * 8u), 256, 0, st>>>(tmp, count, n_buckets, keys);
    if (decode32) k_decode_events<<<blocks, kThreads, 0, st>>>(nt, keys, ctx->sharded ? nullptr : ctx->ids.p, ctx->ev_out.p, ctx->last_add_mode ? 1 : 0);
    ctx->ev_out_valid = decode32;
    ctx->sorted_keys = keys;
    ctx->launches += decode32 ? 6 : 5;
    CU(cudaGetLastError());
    return CB200_OK;
}

// Recursive bucket sort, then the LSD radix sort: for steps with many events at one later time.
static int sort_events_slow(cb200_ctx* ctx, uint32_t nt, uint32_t n_pairs, uint64_t n_solitaire) {
    int rc = build_sort_keys(ctx, n_pairs, n_solitaire, nullptr);
    if (rc != CB200_OK) return rc;
    const uint32_t n_blocks = (nt + kRsTile - 1) / kRsTile;
    CU(ctx->rs_hist.reserve((size_t)16 * n_blocks));
    Key128 *in = ctx->keys_a.p, *outp = ctx->keys_b.p;
    bool sorted = false;
    if ((rc = sort_range(ctx, in, outp, nt, 1, 0, &sorted)) != CB200_OK) return rc;
    if (!sorted) {
        // fallback: LSD radix sort on the digits that differ
        CU(cudaMemsetAsync(&ctx->d_ctr->diff_hi, 0, 16, ctx->stream));
        k_key_diff<<<std::min<uint32_t>(n_blocks, 1024), kThreads, 0, ctx->stream>>>(nt, ctx->keys_a.p, ctx->d_ctr);
        ctx->launches += 1;
        unsigned long long diff[2];
        CU(cudaMemcpyAsync(diff, &ctx->d_ctr->diff_hi, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        const unsigned long long diff_hi = diff[0], diff_lo = diff[1];
        for (int shift = 0; shift < 128; shift += 4) {
            const unsigned long long d = shift < 64 ? (diff_lo >> shift) & 15ull : (diff_hi >> (shift - 64)) & 15ull;
            if (!d) continue;
            k_rs_hist<<<n_blocks, kRsThreads, 0, ctx->stream>>>(in, nt, shift, ctx->rs_hist.p, n_blocks);
            k_rs_scan<<<1, 1024, 0, ctx->stream>>>(ctx->rs_hist.p, 16 * n_blocks);
            k_rs_scatter<<<n_blocks, kRsThreads, 0, ctx->stream>>>(in, outp, nt, shift, ctx->rs_hist.p, n_blocks);
            ctx->launches += 3;
            std::swap(in, outp);
        }
    }
    k_decode_events<<<(nt + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(nt, in, ctx->sharded ? nullptr : ctx->ids.p, ctx->ev_out.p, ctx->last_add_mode ? 1 : 0);
    ctx->sorted_keys = in;
    ctx->ev_out_valid = true;
    ctx->launches += 1;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return CB200_OK;
}

// Sorted events of the last step in ev_out.  `copy_to` (optional): `take` records from `offset` are copied to the host
// behind the sort, so a fetch costs one synchronisation.
static int sort_events_to(cb200_ctx* ctx, EventOut* copy_to, uint64_t offset, uint64_t cap, uint64_t* taken) {
    if (taken) *taken = 0;
    if (ctx->fetch_pending) {  // (a pipelined download may still be sorting on the copy stream: it shares the key buffers)
        int rc = flush_deferred_fetch(ctx);
        if (rc != CB200_OK) return rc;
        CU(cudaStreamSynchronize(ctx->d2h_stream));
    }
    if (ctx->ev_sorted && !ctx->ev_out_valid && ctx->n_ev_sorted) {
        // sorted by a pipelined fetch, which only produced the compact records: decode the 32-byte ones now
        if (ctx->d2h_stream) CU(cudaStreamSynchronize(ctx->d2h_stream));
        if (ctx->sort_oneshot && ctx->h_sort_big[0] != 0) {
            ctx->ev_sorted = false;  // (that sort met an oversized bucket: sort again below, with the fallback)
        } else {
            const uint32_t n = (uint32_t)ctx->n_ev_sorted;
            CU(ctx->ev_out.reserve(n));
            k_decode_events<<<(n + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(n, ctx->sorted_keys, ctx->sharded ? nullptr : ctx->ids.p, ctx->ev_out.p, ctx->last_add_mode ? 1 : 0);
            ctx->launches += 1;
            ctx->ev_out_valid = true;
        }
    }
    const bool was_sorted = ctx->ev_sorted;
    uint32_t nt = 0, n_pairs = 0;
    uint64_t n_solitaire = 0;
    if (!was_sorted) {
        const Counters& c = ctx->last.ctr;
        n_solitaire = c.n_solitaire;
        const uint64_t total = c.n_pair_events + c.n_solitaire;
        ctx->n_ev_sorted = total;
        if (total >= 0xffffffffull) return fail(ctx, CB200_E_STATE, "too many events to sort");
        nt = (uint32_t)total;
        n_pairs = (uint32_t)c.n_pair_events;
        if (nt) {
            CU(ctx->keys_a.reserve(nt)); CU(ctx->keys_b.reserve(nt)); CU(ctx->ev_out.reserve(nt));
            ctx->h_sort_big[0] = 0;
            int rc = ctx->sort_oneshot ? enqueue_sort_oneshot(ctx, nt, n_pairs, n_solitaire, ctx->stream, true) : sort_events_slow(ctx, nt, n_pairs, n_solitaire);
            if (rc != CB200_OK) return rc;
        }
    }
    const uint64_t avail = ctx->n_ev_sorted > offset ? ctx->n_ev_sorted - offset : 0;
    const uint64_t take = copy_to ? std::min(avail, cap) : 0;
    if (take) CU(cudaMemcpyAsync(copy_to, ctx->ev_out.p + offset, take * sizeof(EventOut), cudaMemcpyDeviceToHost, ctx->stream));
    if (take || (!was_sorted && nt)) CU(cudaStreamSynchronize(ctx->stream));
    if (!was_sorted && nt && ctx->sort_oneshot && ctx->h_sort_big[0] != 0) {
        ctx->sort_fallbacks += 1;
        if (getenv("CB200_VERBOSE")) fprintf(stderr, "[cb200] one-shot event sort: %u oversized buckets of %u keys, falling back\n", ctx->h_sort_big[0], nt);
        int rc = sort_events_slow(ctx, nt, n_pairs, n_solitaire);
        if (rc != CB200_OK) return rc;
        if (take) CU(cudaMemcpy(copy_to, ctx->ev_out.p + offset, take * sizeof(EventOut), cudaMemcpyDeviceToHost));
    }
    ctx->ev_sorted = true;
    if (taken) *taken = take;
    return CB200_OK;
}
static int sort_events(cb200_ctx* ctx) { return sort_events_to(ctx, nullptr, 0, 0, nullptr); }

int cb200_fetch_events(cb200_ctx* ctx, cb200_event* buf, uint64_t cap, uint64_t offset, uint64_t* n_out) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->have_step) return fail(ctx, CB200_E_STATE, "fetch_events before bulk_update");
    CU(cudaSetDevice(ctx->device));
    static_assert(sizeof(cb200_event) == sizeof(EventOut), "event layout");
    if (!buf && cap) {
        // (the count is not known before the sort: a null buffer is an error only if something would be written)
        int rc = sort_events(ctx);
        if (rc != CB200_OK) return rc;
        if (ctx->n_ev_sorted > offset) return fail(ctx, CB200_E_ARG, "null event buffer");
        if (n_out) *n_out = 0;
        return CB200_OK;
    }
    uint64_t take = 0;
    int rc = sort_events_to(ctx, reinterpret_cast<EventOut*>(buf), offset, cap, &take);
    if (rc != CB200_OK) return rc;
    if (n_out) *n_out = take;
    return CB200_OK;
}

// ---- pipelined I/O -----------------------------------------------------------------------------------------
static int ensure_io_streams(cb200_ctx* ctx) {
    if (ctx->d2h_stream) return CB200_OK;
    CU(cudaStreamCreateWithPriority(&ctx->d2h_stream, cudaStreamNonBlocking, ctx->prio_low));
    CU(cudaStreamCreateWithPriority(&ctx->h2d_stream, cudaStreamNonBlocking, ctx->prio_low));
    CU(cudaEventCreateWithFlags(&ctx->ev_sorted_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_keys_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_staged_done[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_staged_done[1], cudaEventDisableTiming));
    return CB200_OK;
}

// New velocities for a coming bulk step, copied host -> device on the context's upload stream while the current step (or
// the event download of the previous one) is running.  Up to two sets can be staged; each cb200_bulk_update called with
// vel == NULL consumes the oldest one.
int cb200_stage_velocities(cb200_ctx* ctx, const cb200_vel_view* vel) {
    if (!ctx || !vel) return CB200_E_ARG;
    if (!ctx->have_state) return fail(ctx, CB200_E_STATE, "stage_velocities before upload_state");
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_io_streams(ctx);
    if (rc != CB200_OK) return rc;
    if (ctx->staged_count >= 2) return fail(ctx, CB200_E_STATE, "two velocity sets are staged already: run a bulk step (vel == NULL) first");
    const uint32_t n = ctx->n;
    const int slot = (ctx->staged_head + ctx->staged_count) & 1;
    const double* src[5] = {vel->vel_x, vel->vel_y, vel->rsz_x, vel->rsz_y, vel->end_time};
    for (int k = 0; k < 5; ++k) {
        ctx->staged_has[slot][k] = src[k] != nullptr;
        if (!src[k] || !n) continue;
        CU(ctx->nvel_staged[slot][k].reserve(n));
        CU(cudaMemcpyAsync(ctx->nvel_staged[slot][k].p, src[k], (size_t)n * 8, cudaMemcpyHostToDevice, ctx->h2d_stream));
    }
    CU(cudaEventRecord(ctx->ev_staged_done[slot], ctx->h2d_stream));
    ctx->staged_count += 1;
    return CB200_OK;
}

// The pipelined fetch's sort as two replayed CUDA graphs (sort.cuh "sizes in DEVICE memory"): the order keys on the context's
// stream, then — beside whatever that stream does next — bucket sort and decode into `dst` on the copy stream.
static int capture_to_exec(cb200_ctx* ctx, cudaStream_t st, cudaGraphExec_t* exec, const std::function<void()>& body) {
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGetLastError();
        return CB200_E_CUDA;
    }
    body();
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (e != cudaSuccess || !graph) {
        cudaGetLastError();
        return CB200_E_CUDA;
    }
    e = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return CB200_E_CUDA;
    }
    return CB200_OK;
}
// Second half of a pipelined fetch: the sort graph and the download, on the copy stream.  Called once the next step has been
// queued (cb200_bulk_update / cb200_shard_step / cb200_shard_finish) or when the caller waits for the fetch / reads the queue.
static int flush_deferred_fetch(cb200_ctx* ctx) {
    if (!ctx->fetch_deferred_sort) return CB200_OK;
    cudaGraphExec_t g = ctx->fetch_deferred_sort;
    ctx->fetch_deferred_sort = nullptr;
    CU(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_keys_done, 0));
    CU(cudaGraphLaunch(g, ctx->d2h_stream));
    if (ctx->fetch_take) CU(cudaMemcpyAsync(ctx->fetch_dst, ctx->fetch_deferred_src, ctx->fetch_take * sizeof(Event16), cudaMemcpyDeviceToHost, ctx->d2h_stream));
    return CB200_OK;
}

static int enqueue_sort_graph(cb200_ctx* ctx, uint32_t n_pairs, uint64_t n_solitaire, Event16* dst, size_t cap_keys) {
    int log2_cap = 6;
    while ((1ull << log2_cap) * 8ull < cap_keys && log2_cap < 24) ++log2_cap;
    const uint32_t nb_max = 2u << log2_cap;
    const uint32_t scan_blocks = (nb_max + kBs2Chunk - 1) / kBs2Chunk;
    DevBuf<uint32_t>& tab = ctx->bs_tab[0];
    CU(tab.reserve(2 * (size_t)nb_max + 16 + scan_blocks));
    CU(ctx->sort_dyn.reserve(1));
    uint32_t* count = tab.p;                // [n_buckets + 1] then: offsets
    uint32_t* cursor = tab.p + nb_max + 4;  // [n_buckets]
    BsRange* range = reinterpret_cast<BsRange*>(tab.p + 2 * (size_t)nb_max + 4);  // [2]: time keys, first labels
    uint32_t* block_total = tab.p + 2 * (size_t)nb_max + 12;
    Key128 *keys = ctx->keys_a.p, *tmp = ctx->keys_b.p;
    SortDyn* dyn = ctx->sort_dyn.p;
    const uint32_t n = ctx->n;
    const uint32_t g_rows = std::max<uint32_t>(1, std::min<uint32_t>((n + kThreads - 1) / kThreads, (uint32_t)ctx->sm_count * 8u));
    const uint32_t g_keys = (uint32_t)std::max<size_t>(1, std::min<size_t>((cap_keys + kThreads - 1) / kThreads, (size_t)ctx->sm_count * 16u));
    // signature of everything the captured launches carry
    uint64_t h = 0xcbf29ce484222325ull;
    auto mix = [&h](uint64_t v) {
        h ^= v;
        h *= 0x100000001b3ull;
        h ^= h >> 29;
    };
    const void* ptrs[] = {count, keys, tmp, dyn, dst, ctx->events.p, ctx->reit.p, ctx->col[9].p, ctx->label.p, ctx->d_ctr, ctx->d_sort_big};
    for (const void* q : ptrs) mix((uint64_t)(uintptr_t)q);
    const uint32_t cap_events = (uint32_t)std::min<size_t>(ctx->events.cap, 0xffffffffu), add_mode = ctx->last_add_mode ? 1u : 0u;
    mix(n); mix(nb_max); mix(g_keys); mix((uint64_t)log2_cap); mix(cap_events); mix(add_mode);
    if (!h) h = 1;
    cb200_ctx::FetchGraph* slot = nullptr;
    for (auto& g : ctx->fetch_graphs)
        if (g.sig == h && g.keys && g.sort) slot = &g;
    if (!slot) {
        if (getenv("CB200_VERBOSE")) fprintf(stderr, "[cb200] capturing the event-sort graphs (sig %016llx, n %u, cap_keys %zu)\n", (unsigned long long)h, n, cap_keys);
        cb200_ctx::FetchGraph& g = ctx->fetch_graphs[ctx->fetch_graph_next];
        ctx->fetch_graph_next = (ctx->fetch_graph_next + 1) % cb200_ctx::kFetchGraphs;
        if (g.keys) cudaGraphExecDestroy(g.keys);
        if (g.sort) cudaGraphExecDestroy(g.sort);
        g = cb200_ctx::FetchGraph{};
        int rc = capture_to_exec(ctx, ctx->stream, &g.keys, [&]() {
            k_sortdyn_init<<<1, 1, 0, ctx->stream>>>(ctx->d_ctr, cap_events, add_mode, dyn, range, log2_cap);
            k_keys_all_dyn<<<g_rows + g_keys, kThreads, 0, ctx->stream>>>(dyn, g_rows, n, ctx->reit.p, ctx->col[9].p, ctx->label.p, ctx->events.p, keys, ctx->d_ctr, range,
                                                                         range + 1);
        });
        if (rc != CB200_OK) return fail(ctx, rc, "capture of the event-key graph failed");
        cudaStream_t st = ctx->d2h_stream;
        rc = capture_to_exec(ctx, st, &g.sort, [&]() {
            k_sortdyn_zero<<<(uint32_t)ctx->sm_count * 2u, kThreads, 0, st>>>(dyn, count);
            k_bs2_count_dyn<<<g_keys, kThreads, 0, st>>>(dyn, keys, range, count);
            k_bs2_scan_local_dyn<<<scan_blocks, 256, 0, st>>>(count, dyn, cursor, block_total);
            k_bs2_scan_add_dyn<<<scan_blocks, 256, 0, st>>>(count, dyn, cursor, block_total, ctx->d_sort_big);
            k_bs2_scatter_dyn<<<g_keys, kThreads, 0, st>>>(dyn, keys, range, cursor, tmp);
            k_bs_local_dyn<<<(uint32_t)ctx->sm_count * 8u, 256, 0, st>>>(tmp, count, dyn, keys);
            k_decode_events16_dyn<<<g_keys, kThreads, 0, st>>>(dyn, keys, dst);
        });
        if (rc != CB200_OK) return fail(ctx, rc, "capture of the event-sort graph failed");
        g.sig = h;
        slot = &g;
    }
    (void)n_pairs; (void)n_solitaire;  // (the graphs read the counts where the step left them: Counters on the device)
    CU(cudaGraphLaunch(slot->keys, ctx->stream));
    CU(cudaEventRecord(ctx->ev_keys_done, ctx->stream));
    ctx->fetch_deferred_sort = slot->sort;  // (launched by flush_deferred_fetch)
    ctx->ev_out_valid = false;
    ctx->sorted_keys = keys;
    ctx->launches += 9;
    return CB200_OK;
}

// The sorted events of the last step in the compact form (cb200_event16), downloaded on the context's copy stream: returns
// as soon as the sort and the copy are queued, so the caller can start the next bulk step; cb200_fetch_events_end waits for
// the copy.  *n_out = events that will be in buf[0 .. *n_out) (at most cap).
int cb200_fetch_events_begin(cb200_ctx* ctx, cb200_event16* buf, uint64_t cap, uint64_t* n_out) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->have_step) return fail(ctx, CB200_E_STATE, "fetch_events before bulk_update");
    if (ctx->fetch_pending) return fail(ctx, CB200_E_STATE, "cb200_fetch_events_end the pending download first");
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_io_streams(ctx);
    if (rc != CB200_OK) return rc;
    static_assert(sizeof(cb200_event16) == sizeof(Event16), "compact event layout");
    const Counters& c = ctx->last.ctr;
    const uint64_t total = c.n_pair_events + c.n_solitaire;
    if (total >= 0xffffffffull) return fail(ctx, CB200_E_STATE, "too many events to sort");
    const uint32_t nt = (uint32_t)total;
    const uint64_t take = buf ? std::min<uint64_t>(total, cap) : 0;
    if (n_out) *n_out = take;
    ctx->ev16_parity ^= 1;
    DevBuf<Event16>& dst = ctx->ev16[ctx->ev16_parity];
    ctx->fetch_oneshot_pending = false;
    if (nt) {
        // (sized for everything a step can queue, so that the buffers — and the graphs that carry their addresses — stay put)
        const size_t cap_keys = std::max<size_t>(nt, ctx->events.cap + (size_t)ctx->n);
        CU(ctx->keys_a.reserve(cap_keys)); CU(ctx->keys_b.reserve(cap_keys)); CU(dst.reserve(cap_keys));
        bool on_copy_stream = false, decoded = false;
        if (!ctx->ev_sorted) {
            ctx->n_ev_sorted = total;
            ctx->h_sort_big[0] = 0;
            if (ctx->sort_oneshot && ctx->fetch_graph_enabled && !ctx->capturing) {
                if ((rc = enqueue_sort_graph(ctx, (uint32_t)c.n_pair_events, c.n_solitaire, dst.p, cap_keys)) != CB200_OK) return rc;
                ctx->fetch_oneshot_pending = true;
                on_copy_stream = decoded = true;
            } else if (ctx->sort_oneshot) {
                // keys on the context's stream (a few microseconds), the bucket sort on the copy stream: the next bulk step does
                // not wait for it
                if ((rc = enqueue_sort_oneshot(ctx, nt, (uint32_t)c.n_pair_events, c.n_solitaire, ctx->d2h_stream, false)) != CB200_OK) return rc;
                ctx->fetch_oneshot_pending = true;  // (whether it held is known once the stream has run: checked in _end)
                on_copy_stream = true;
            } else {
                CU(ctx->ev_out.reserve(nt));
                if ((rc = sort_events_slow(ctx, nt, (uint32_t)c.n_pair_events, c.n_solitaire)) != CB200_OK) return rc;
            }
            ctx->ev_sorted = true;
        }
        // the sorted keys -> compact records, on the stream that sorted them
        cudaStream_t ds = on_copy_stream ? ctx->d2h_stream : ctx->stream;
        if (!decoded) {
            k_decode_events16<<<(nt + kThreads - 1) / kThreads, kThreads, 0, ds>>>(nt, ctx->sorted_keys, dst.p, ctx->last_add_mode ? 1 : 0);
            ctx->launches += 1;
        }
        if (!on_copy_stream) {
            CU(cudaEventRecord(ctx->ev_sorted_done, ctx->stream));
            CU(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_sorted_done, 0));
        }
    }
    ctx->fetch_dst = reinterpret_cast<Event16*>(buf);
    ctx->fetch_take = take;
    ctx->fetch_deferred_src = dst.p;
    if (!ctx->fetch_deferred_sort && take) CU(cudaMemcpyAsync(buf, dst.p, take * sizeof(Event16), cudaMemcpyDeviceToHost, ctx->d2h_stream));
    ctx->fetch_pending = true;
    ctx->fetch_nt = nt;
    ctx->fetch_np = (uint32_t)c.n_pair_events;
    ctx->fetch_ns = c.n_solitaire;
    return CB200_OK;
}

int cb200_fetch_events_end(cb200_ctx* ctx) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->fetch_pending) return CB200_OK;
    CU(cudaSetDevice(ctx->device));
    {
        int rc = flush_deferred_fetch(ctx);
        if (rc != CB200_OK) return rc;
    }
    CU(cudaStreamSynchronize(ctx->d2h_stream));
    ctx->fetch_pending = false;
    if (ctx->fetch_oneshot_pending && ctx->h_sort_big[0] != 0) {
        // the one-shot sort met an oversized bucket: the download holds a partially sorted list.  This can only be repaired
        // while the step's events are still the last step's — the caller must not have started another step
        ctx->sort_fallbacks += 1;
        return fail(ctx, CB200_E_STATE, "the one-shot event sort met an oversized bucket (a heavy tie at a later time): fetch this step's events with cb200_fetch_events instead");
    }
    return CB200_OK;
}

// gid -> HbId for the debug views (one device: the id column; sharded: identity).
static int gid_to_id_table(cb200_ctx* ctx, std::vector<uint64_t>& ids) {
    ids.clear();
    if (ctx->sharded) return CB200_OK;
    ids.resize(ctx->n);
    if (ctx->n) CU(cudaMemcpy(ids.data(), ctx->ids.p, (size_t)ctx->n * 8, cudaMemcpyDeviceToHost));
    return CB200_OK;
}
static inline uint64_t id_of(const std::vector<uint64_t>& ids, uint32_t gid) { return ids.empty() ? (uint64_t)gid : ids[gid]; }

int cb200_fetch_candidate_pairs(cb200_ctx* ctx, uint64_t* id_hi, uint64_t* id_lo, uint64_t cap, uint64_t* n_out) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->have_step) return fail(ctx, CB200_E_STATE, "fetch_candidate_pairs before bulk_update");
    CU(cudaSetDevice(ctx->device));
    const uint64_t total = ctx->last.ctr.n_collide;
    if (n_out) *n_out = total;
    if (cap == 0 || total == 0) return CB200_OK;
    std::vector<uint64_t> ids;
    int rc = gid_to_id_table(ctx, ids);
    if (rc != CB200_OK) return rc;
    // the step keeps (entry, predecessor) work items; re-run the candidate test over them and list what passes
    const ShardGeom sh = ctx->sharded ? ctx->shard : whole_world();
    CandArgs cd{};
    cd.geom = ctx->geom;
    cd.col_lo = sh.col_lo;
    cd.col_hi = sh.col_hi;
    cd.ov = OverlapSet{reinterpret_cast<const ulonglong2*>(ctx->ov_table.p), ctx->ov_bits.p, ctx->ov_bucket_mask, ctx->n_ov, ctx->ov_pair_bits.p, ctx->ov_pair_mask};
    DevBuf<uint2> out;
    DevBuf<unsigned long long> cnt;
    CU(out.reserve(total));
    CU(cnt.reserve(1));
    CU(cudaMemsetAsync(cnt.p, 0, 8, ctx->stream));
    if (ctx->last_step_tile) {
        // the tile kernel again, listing instead of solving (state, visitor slabs and this step's counters are untouched)
        const VisSpace vs{ctx->n, ctx->n, 0u, 1, 0};
        TileSolveArgs ts{};
        ts.sv = solve_args_of(ctx, sh, vs);
        ts.s = cols_of(ctx);
        ts.label = ctx->label.p;
        ts.bink = ctx->bink.p;
        ts.tg = ctx->tgeom;
        ts.tt = TileTables{ctx->tile_rows.p, ctx->home.p, ctx->vis_count.p, ctx->vis.p, ctx->tile_vcap, ctx->tgeom.n_tiles};
        ts.cap_rec = ctx->tile_cap_rec;
        ts.cap_ent = ctx->tile_cap_ent;
        // every pair that passes candidate_geom becomes a listed survivor (no swept-bounds filter): the list may need more room
        for (int attempt = 0; attempt < 3; ++attempt) {
            ts.surv = SurvList{ctx->surv.p, (uint32_t)std::min<size_t>(ctx->surv.cap, 0x7fffffffu)};
            CU(cudaMemsetAsync(&ctx->d_ctr->n_surv, 0, 8, ctx->stream));
            k_tile_bin<true><<<ctx->grid_t, kTileThreads, ctx->tile_smem, ctx->stream>>>(ts);
            unsigned long long n_surv = 0;
            CU(cudaMemcpyAsync(&n_surv, &ctx->d_ctr->n_surv, 8, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            if (n_surv <= ts.surv.cap) break;
            CU(ctx->surv.reserve((size_t)n_surv + 1024));
        }
        k_tile_finish<true><<<ctx->grid_s, kThreads, 0, ctx->stream>>>(ts, out.p, cnt.p, total);
    } else
    k_list_candidates<<<ctx->sm_count * 4, kThreads, 0, ctx->stream>>>(cd, work_lists_of(ctx), ctx->rec.p, out.p, cnt.p, total);
    if (!ctx->last_step_tile && ctx->last_dense_rank != kDenseOff) {
        DenseArgs da{};
        da.list = dense_list_of(ctx);
        da.entries = reinterpret_cast<const uint2*>(ctx->entries.p);
        da.cap_entries = (uint32_t)std::min<size_t>(ctx->entries.cap, 0xffffffffu);
        k_list_candidates_dense<<<ctx->sm_count * 4, kThreads, 0, ctx->stream>>>(cd, da, ctx->d_ctr, ctx->rec.p, out.p, cnt.p, total);
    }
    ctx->launches += 2;
    std::vector<uint2> h(total);
    unsigned long long got = 0;
    CU(cudaMemcpyAsync(h.data(), out.p, total * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&got, cnt.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    out.release();
    cnt.release();
    if (got != total) return fail(ctx, CB200_E_STATE, "candidate listing disagrees with the step's count");
    for (uint64_t i = 0; i < total && i < cap; ++i) {
        if (id_hi) id_hi[i] = id_of(ids, h[i].x);
        if (id_lo) id_lo[i] = id_of(ids, h[i].y);
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
    {
        int rc = rebuild_global_grid(ctx);  // (after a tile step)
        if (rc != CB200_OK) return rc;
    }
    std::vector<CellEntry> h(total);
    std::vector<uint64_t> ids;
    int rc = gid_to_id_table(ctx, ids);
    if (rc != CB200_OK) return rc;
    CU(cudaMemcpy(h.data(), ctx->entries.p, total * sizeof(CellEntry), cudaMemcpyDeviceToHost));
    // an entry names its record; the record's first 16 bytes hold sx, sy, span, gid
    const uint32_t n_rec = ctx->sharded ? (uint32_t)std::min<size_t>(ctx->rec.cap, 0xffffffffu) : ctx->n;
    std::vector<int32_t> heads((size_t)n_rec * 4);
    if (n_rec) CU(cudaMemcpy2D(heads.data(), 16, ctx->rec.p, sizeof(HbRec), 16, n_rec, cudaMemcpyDeviceToHost));
    const uint32_t mw = (1u << ctx->geom.lw) - 1u, mh = (1u << ctx->geom.lh) - 1u;
    for (uint64_t i = 0; i < total && i < cap; ++i) {
        const int32_t* q = heads.data() + 4 * (size_t)h[i].vidx;
        uint32_t sxs, sys;
        ctx->geom.unslot(h[i].key & kSlotMask, &sxs, &sys);
        // the unique cell of the rect that folds onto this slot
        const int32_t x = q[0] + (int32_t)((sxs - (uint32_t)q[0]) & mw), y = q[1] + (int32_t)((sys - (uint32_t)q[1]) & mh);
        if (cell_x) cell_x[i] = x;
        if (cell_y) cell_y[i] = y;
        if (id) id[i] = id_of(ids, (uint32_t)q[3]);
    }
    return CB200_OK;
}

int cb200_fetch_index_rects(cb200_ctx* ctx, int32_t* rects, uint64_t cap_hitboxes) {
    if (!ctx || !rects) return CB200_E_ARG;
    if (!ctx->have_step) return fail(ctx, CB200_E_STATE, "fetch_index_rects before bulk_update");
    if (ctx->sharded) return fail(ctx, CB200_E_STATE, "fetch_index_rects is a one-device debug view");
    if (cap_hitboxes < ctx->n) return fail(ctx, CB200_E_ARG, "rect buffer too small");
    CU(cudaSetDevice(ctx->device));
    {
        int rc = rebuild_global_grid(ctx);  // (after a tile step the records are not all there)
        if (rc != CB200_OK) return rc;
    }
    // strided D2H: first 16 bytes of every 96-byte record = sx, sy, span, label -> decode the span, place by label
    std::vector<int32_t> raw((size_t)ctx->n * 4);
    if (ctx->n) CU(cudaMemcpy2D(raw.data(), 16, ctx->rec.p, sizeof(HbRec), 16, ctx->n, cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < ctx->n; ++i) {
        const int32_t* q = raw.data() + 4 * i;
        const uint32_t span = (uint32_t)q[2], label = (uint32_t)q[3];
        int32_t* r = rects + 4 * (size_t)label;
        r[0] = q[0];
        r[1] = q[1];
        r[2] = q[0] + (int32_t)(span & 0xffffu);
        r[3] = q[1] + (int32_t)(span >> 16);
    }
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
int cb200_last_step_path(const cb200_ctx* ctx) { return ctx && ctx->have_step ? (ctx->last_step_tile ? 1 : 0) : -1; }
uint64_t cb200_tile_fallbacks(const cb200_ctx* ctx) { return ctx ? ctx->tile_fallbacks : 0; }
int cb200_get_tile(const cb200_ctx* ctx, int* log2_slots, int* cells_x, int* cells_y) {
    if (!ctx) return CB200_E_ARG;
    if (log2_slots) *log2_slots = ctx->tile_ok ? ctx->tgeom.tj : 0;
    if (cells_x) *cells_x = ctx->tile_ok ? 1 << ctx->tgeom.bx : 0;
    if (cells_y) *cells_y = ctx->tile_ok ? 1 << ctx->tgeom.by : 0;
    return CB200_OK;
}

int cb200_last_step_timing(cb200_ctx* ctx, float* total_ms, float* stage_ms, int n_stages) {
    if (!ctx) return CB200_E_ARG;
    if (!ctx->timing_valid || ctx->timed_steps == 0) return fail(ctx, CB200_E_STATE, "no timed step yet");
    const double k = 1.0 / (double)ctx->timed_steps;
    if (total_ms) *total_ms = (float)(ctx->total_ms_sum * k);
    for (int s = 0; s < n_stages && s < ST_COUNT; ++s)
        if (stage_ms) stage_ms[s] = (float)(ctx->stage_ms_sum[s] * k);
    return CB200_OK;
}
int cb200_set_stage_timing(cb200_ctx* ctx, int on) {
    if (!ctx) return CB200_E_ARG;
    ctx->stage_timing = on != 0;
    return cb200_reset_timing(ctx);
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

#include "collider.inl"
